"""Sample storage for the RWR loop (``pipeline/sample.py`` writes, ``pipeline/finetune.py`` reads) -- the role of the
reference's HDF5 shard files (``ddpo/utils/hdf5.py``: ``H5Writer`` :72-204, ``RemoteWriter`` :245-349, ``RemoteReader``
:352-461) with the same field schema and call surface (``configure(field, encode_fn=, decode_fn=)``,
``add_batch(batch, mask=)``, ``close()``; reader ``len``, ``[idx]``, ``get(idx, field)``, ``make_weights(field,
temperature, by_prompt)``), the same codecs (JPEG q=95 images :25-44, pickled generic objects :47-53) and the same
split rule (a new shard every ``split_size`` samples, names ``<worker>_<index>``).

Container: ``h5py`` / ``gcsfs`` are not installable offline and the reference's writer requires a ``gs://`` bucket
(:249-259), so shards are ``.npz`` files on the local mirror of the path (``utils.fs.localize``); every field is stored
as an object array of per-sample encoded values -- a documented deviation of the container, not of the schema."""
import io
import os
import pickle

import numpy as np

from . import filesystem


def encode_jpeg(x, quality=95):
    from PIL import Image
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating):
        assert np.abs(x).max() <= 1.0
        x = (x * 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(x).save(buf, "JPEG", quality=quality)
    return np.frombuffer(buf.getvalue(), dtype=np.uint8)


def decode_jpeg(jpeg):
    from PIL import Image
    return np.array(Image.open(io.BytesIO(np.asarray(jpeg, np.uint8).tobytes()))) / 255.0


def encode_generic(x):
    return np.frombuffer(pickle.dumps(x), dtype=np.uint8)


def decode_generic(x):
    return pickle.loads(np.asarray(x, np.uint8).tobytes())


def softmax_ref(x, temperature=1.0):
    """``utils.softmax_ref`` (reference ``ddpo/utils/array.py:32-41``): softmax(x * temperature) of a 1-D array."""
    x = np.asarray(x)
    assert x.ndim == 1
    z = x * temperature
    z = z - z.max()
    e = np.exp(z)
    return e / e.sum()


_CODECS = {"jpeg": (encode_jpeg, decode_jpeg), "generic": (encode_generic, decode_generic), "raw": (None, None)}


class ShardWriter:
    def __init__(self, savepath, split_size=1600, worker=None):
        from ..training import distributed
        self.savepath = filesystem.localize(savepath)
        os.makedirs(self.savepath, exist_ok=True)
        self.split_size = int(split_size)
        self.worker = distributed.rank() if worker is None else worker
        self._codec = {}
        self._data = {}
        self._index = 0
        self._total = 0

    def __len__(self):
        return self._total

    def configure(self, field, encode_fn=None, decode_fn=None, **kwargs):
        name = "raw"
        for k, (e, _) in _CODECS.items():
            if encode_fn is e and e is not None:
                name = k
        if encode_fn is not None and name == "raw":
            raise ValueError("only utils.encode_jpeg / utils.encode_generic codecs can be recorded in a shard")
        self._codec[field] = name
        self._data.setdefault(field, [])

    def add_batch(self, batch, mask=None, **kwargs):
        sizes = [len(v) for v in batch.values()]
        assert len(set(sizes)) == 1, f"Batch sizes must be equal, got {sizes}"
        indices = range(sizes[0]) if mask is None else np.where(np.asarray(mask).reshape(-1))[0]
        print(f"[ utils/shards ] Adding {len(indices)} samples | [{self._total}, {self._total + len(indices)}]")
        for i in indices:
            for key, val in batch.items():
                if key not in self._codec:
                    self.configure(key)
                enc = _CODECS[self._codec[key]][0]
                self._data[key].append(enc(val[i]) if enc is not None else np.asarray(val[i]))
            self._total += 1
            if len(next(iter(self._data.values()))) >= self.split_size:
                self._flush()
        return len(indices)

    def _flush(self):
        n = len(next(iter(self._data.values()))) if self._data else 0
        if n == 0:
            return
        path = os.path.join(self.savepath, f"{self.worker}_{self._index}.npz")
        arrays = {}
        for k, v in self._data.items():
            a = np.empty(len(v), dtype=object)
            for i, x in enumerate(v):
                a[i] = x
            arrays[k] = a
        arrays["__codecs__"] = np.array([pickle.dumps(self._codec)], dtype=object)
        with open(path, "wb") as f:
            np.savez(f, **arrays)
        print(f"[ utils/shards ] Wrote {n} samples to {path}")
        self._data = {k: [] for k in self._data}
        self._index += 1

    def close(self):
        self._flush()


class ShardReader:
    def __init__(self, loadpath):
        self.loadpath = filesystem.localize(loadpath)
        files = sorted(f for f in os.listdir(self.loadpath) if f.endswith(".npz"))
        if not files:
            raise FileNotFoundError(f"no sample shards under {self.loadpath}")
        self._fields, self._codec = {}, {}
        for f in files:
            with np.load(os.path.join(self.loadpath, f), allow_pickle=True) as z:
                codec = pickle.loads(z["__codecs__"][0])
                for k in z.files:
                    if k == "__codecs__":
                        continue
                    self._fields.setdefault(k, []).extend(list(z[k]))
                    self._codec[k] = codec[k]
        self._keys = sorted(self._fields)
        self._len = len(self._fields[self._keys[0]])
        assert all(len(v) == self._len for v in self._fields.values())
        self.weighted = False
        print(f"[ utils/shards ] Found {self._len} samples in {len(files)} shards under {self.loadpath}")

    def __len__(self):
        return self._len

    def sizes(self):
        return {k: len(v) for k, v in self._fields.items()}

    def get(self, idx, field="images"):
        dec = _CODECS[self._codec[field]][1]
        one = lambda x: dec(x) if dec is not None else x
        if isinstance(idx, slice):
            return np.array([one(x) for x in self._fields[field][idx]])
        return one(self._fields[field][int(idx)])

    def __getitem__(self, idx):
        batch = {k: self.get(idx, field=k) for k in self._keys}
        if self.weighted:
            batch["weights"] = self.weights[idx]
        return batch

    def make_weights(self, field, temperature, by_prompt):
        """Reference ``hdf5.py:437-461``: dataset-level softmax(reward * temperature) * N, optionally per prompt."""
        labels = np.asarray(self.get(slice(0, len(self)), field), np.float64).squeeze()
        if by_prompt:
            prompts = np.asarray(self.get(slice(0, len(self)), "inference_prompts")).squeeze()
            self.weights = np.empty_like(labels)
            for prompt in np.unique(prompts):
                mask = prompts == prompt
                self.weights[mask] = softmax_ref(labels[mask], temperature=temperature) * mask.sum()
        else:
            self.weights = softmax_ref(labels, temperature=temperature) * len(self)
        self.weighted = True
        cumsum = np.cumsum(np.sort(self.weights)[::-1] / len(self))
        n = ((cumsum <= 0.9) * np.arange(len(cumsum))).max()
        print(f"[ utils/shards ] Weights sanity check: {n} / {len(cumsum)} samples account for 90% of the weight | "
              f"temperature: {temperature}")
