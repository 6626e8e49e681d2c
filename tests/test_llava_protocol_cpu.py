"""The LLaVA reward's pickle-over-HTTP wire format (reference ddpo/training/callbacks.py:464-537) against a local
stand-in server that checks the request and answers in the reference's response schema."""
import io
import pickle
import threading
from http.server import BaseHTTPRequestHandler, HTTPServer

import numpy as np


class _Handler(BaseHTTPRequestHandler):
    seen = []

    def do_POST(self):
        data = pickle.loads(self.rfile.read(int(self.headers["Content-Length"])))
        from PIL import Image
        n = len(data["images"])
        sizes = [Image.open(io.BytesIO(b)).size for b in data["images"]]
        _Handler.seen.append((n, sizes, data["queries"], data["answers"]))
        body = pickle.dumps({"recall": np.arange(n, dtype=np.float32)[:, None] / 10, "precision": np.ones((n, 1)),
                             "f1": np.full((n, 1), 0.5), "outputs": [[f"caption {i}"] for i in range(n)]})
        self.send_response(200)
        self.send_header("Content-Length", str(len(body)))
        self.end_headers()
        self.wfile.write(body)

    def log_message(self, *a):
        pass


def test_llava_bertscore_wire_format_and_fallback(monkeypatch):
    from ddpo_b200.training import callbacks as C
    srv = HTTPServer(("127.0.0.1", 0), _Handler)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        fn = C.callback_fns["llava_bertscore"](url=f"http://127.0.0.1:{srv.server_address[1]}")
        images = np.random.default_rng(0).random((18, 32, 32, 3)).astype(np.float32)
        prompts = [f"a dog riding bike {i}" for i in range(18)]
        scores, info = fn(images, prompts, [{}] * 18)
    finally:
        srv.shutdown()
    assert scores.shape == (18,)                                   # LLaVA rewards are [N], not [N, 1]
    np.testing.assert_allclose(scores[:9], np.arange(9) / 10, rtol=1e-6)
    assert [s[0] for s in _Handler.seen] == [9, 9]                 # np.array_split(18 images, ceil(18 / 16) = 2 batches)
    n, sizes, queries, answers = _Handler.seen[0]
    assert sizes == [(32, 32)] * 9
    assert queries == [["Answer concisely: what is going on in this image?"]] * 9
    assert answers[3] == ["The image contains a dog riding bike 3"]
    assert info["f1"].shape == (18,) and info["outputs"][10] == "caption 1"
    # no server configured -> the cached-score stub (deterministic per prompt batch)
    monkeypatch.delenv(C.LLAVA_URL_ENV, raising=False)
    monkeypatch.delenv(C.ALLOW_STUB_ENV, raising=False)
    import pytest
    with pytest.raises(RuntimeError):                              # a reward that ignores the images needs an explicit opt-in
        C.callback_fns["llava_bertscore"]()
    monkeypatch.setenv(C.ALLOW_STUB_ENV, "1")
    with pytest.warns(RuntimeWarning):
        stub = C.callback_fns["llava_bertscore"]()
    a, meta = stub(images, prompts, None)
    b, _ = stub(images, prompts, None)
    assert a.shape == (18,) and np.array_equal(a, b) and meta == {"stub": True}
