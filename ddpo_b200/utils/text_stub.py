"""Stand-ins for the CLIP tokenizer / text encoder when no checkpoint files exist (this build has no network:
``transformers`` is importable but neither ``CLIPTokenizer`` vocabularies nor ``FlaxCLIPTextModel`` weights are).

The reference embeds prompts on the host CPU (``pipeline/policy_gradient.py:185-187``) and feeds the U-Net
``[B, 77, D]`` float32 hidden states; everything downstream depends only on that tensor.  ``StubTokenizer`` maps a
prompt to deterministic ids (so that the tokenizer encode -> decode round trip the reference uses to canonicalise
prompts, ``:329-335``, is the identity on normalised text of up to 75 characters, in ANY process) and ``StubTextEncoder`` maps ids to a deterministic
N(0,1) embedding seeded by the ids -- same prompt => same conditioning, different prompts => independent ones."""
import hashlib

import numpy as np


class StubTokenizer:
    """Character-level, STATELESS and reversible: id = 1000 + code point.  Every process decodes every other process's
    ids to the same string (the multi-rank driver all-gathers prompt ids and decodes them locally to key the per-prompt
    reward statistics, reference ``pipeline/policy_gradient.py:329-335``), with no vocabulary to share."""
    model_max_length = 77
    bos, eos = 49406, 49407
    _base, _span = 1000, 40000

    @staticmethod
    def normalise(text):
        return " ".join(str(text).lower().split())

    def __call__(self, prompts, padding="max_length", max_length=None, truncation=True, return_tensors="np"):
        if isinstance(prompts, str):
            prompts = [prompts]
        L = max_length or self.model_max_length
        ids = np.full((len(prompts), L), self.eos, np.int64)
        for r, p in enumerate(prompts):
            toks = [self.bos] + [self._base + min(ord(ch), self._span - 1) for ch in self.normalise(p)][: L - 2] + [self.eos]
            ids[r, : len(toks)] = toks
        return type("Encoding", (), {"input_ids": ids})()

    def batch_decode(self, ids, skip_special_tokens=True):
        out = []
        for row in np.asarray(ids):
            out.append("".join(chr(int(i) - self._base) for i in row
                               if self._base <= int(i) < self._base + self._span).strip())
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
