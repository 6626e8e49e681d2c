"""Stand-ins for the CLIP tokenizer / text encoder when no checkpoint files exist (this build has no network:
``transformers`` is importable but neither ``CLIPTokenizer`` vocabularies nor ``FlaxCLIPTextModel`` weights are).

The reference embeds prompts on the host CPU (``pipeline/policy_gradient.py:185-187``) and feeds the U-Net
``[B, 77, D]`` float32 hidden states; everything downstream depends only on that tensor.  ``StubTokenizer`` maps a
prompt to deterministic ids (so that the tokenizer encode -> decode round trip the reference uses to canonicalise
prompts, ``:329-335``, is the identity on normalised text) and ``StubTextEncoder`` maps ids to a deterministic
N(0,1) embedding seeded by the ids -- same prompt => same conditioning, different prompts => independent ones."""
import hashlib

import numpy as np


class StubTokenizer:
    model_max_length = 77
    bos, eos = 49406, 49407

    def __init__(self):
        self._vocab = {}
        self._inv = {}

    def _id(self, word):
        if word not in self._vocab:
            i = 1000 + int(hashlib.sha1(word.encode()).hexdigest(), 16) % 40000
            while i in self._inv and self._inv[i] != word:
                i += 1
            self._vocab[word] = i
            self._inv[i] = word
        return self._vocab[word]

    def __call__(self, prompts, padding="max_length", max_length=None, truncation=True, return_tensors="np"):
        if isinstance(prompts, str):
            prompts = [prompts]
        L = max_length or self.model_max_length
        ids = np.full((len(prompts), L), self.eos, np.int64)
        for r, p in enumerate(prompts):
            toks = [self.bos] + [self._id(w) for w in str(p).lower().split()][: L - 2] + [self.eos]
            ids[r, : len(toks)] = toks
        return type("Encoding", (), {"input_ids": ids})()

    def batch_decode(self, ids, skip_special_tokens=True):
        out = []
        for row in np.asarray(ids):
            words = [self._inv.get(int(i), "") for i in row if int(i) not in (self.bos, self.eos)]
            out.append(" ".join(w for w in words if w))
        return out


class StubTextEncoder:
    """``text_encoder(input_ids, params=None, train=False)[0]`` -> float32 ``[B, L, dim]``."""

    def __init__(self, dim=1024):
        self.dim = dim
        self.params = {}

    def __call__(self, input_ids, params=None, train=False):
        ids = np.asarray(input_ids)
        out = np.empty((ids.shape[0], ids.shape[1], self.dim), np.float32)
        for r, row in enumerate(ids):
            seed = int(hashlib.sha1(row.astype(np.int64).tobytes()).hexdigest(), 16) % (2 ** 32)
            out[r] = np.random.default_rng(seed).standard_normal((ids.shape[1], self.dim), dtype=np.float32)
        return (out,)
