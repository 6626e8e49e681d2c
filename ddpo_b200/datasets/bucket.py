"""RWR training set -- mirror of the reference's ``ddpo/datasets/bucket.py`` (``BucketDataset`` :10-51,
``select_caption`` :59-63, ``make_uncond_text`` :66-73, ``collate_fn`` :76-115, ``get_bucket_loader`` :118-153) on the
local shard reader.  A plain Python batch iterator stands in for ``torch.utils.data.DataLoader(shuffle=False,
drop_last=True)``: the items are tiny (8 KB of posterior moments + token ids) and go straight to pinned host memory."""
import random

import numpy as np

from .. import utils


class BucketDataset:
    def __init__(self, reader):
        self.reader = reader
        self.transform_fn = lambda x: x
        self._max_size = None
        self._offset = 0
        self._shuffled = np.arange(len(self))

    def __len__(self):
        return self._max_size or len(self.reader)

    def __getitem__(self, idx):
        worker_idx = self._offset + idx
        shuffled_idx = self._shuffled[worker_idx]
        x = self.reader[shuffled_idx]
        return self.transform_fn(x) | {"idx": worker_idx, "shuffled_idx": shuffled_idx}

    def shuffle(self):
        print("[ datasets/bucket ] Shuffling dataset")
        self._shuffled = np.random.permutation(self._shuffled)

    def shard(self):
        from ..training import distributed
        host_id, n_hosts = distributed.rank(), distributed.world_size()
        per = len(self) // n_hosts
        self._max_size, self._offset = per, host_id * per
        print(f"[ datasets/bucket ] Host: {host_id} | Samples per host: {self._max_size} | Offset: {self._offset}")

    def make_weights(self, *args, **kwargs):
        self.reader.make_weights(*args, **kwargs)

    def with_transform(self, transform_fn):
        self.transform_fn = transform_fn

    def subsample(self, N):
        self._max_size = N


def select_caption(examples, field="training_prompts"):
    caption = examples[field]
    if isinstance(caption, (list, tuple, np.ndarray)):
        caption = random.choice(list(caption))
    examples["text"] = caption
    return examples


def make_uncond_text(tokenizer, batch_size):
    return tokenizer([""] * batch_size, padding="max_length", max_length=tokenizer.model_max_length,
                     return_tensors="np").input_ids


def collate_fn(tokenizer, examples, image_field="vae", text_field="input_ids"):
    pixel_values = np.stack([e[image_field] for e in examples]).astype(np.float32)
    captions = [e["text"] for e in examples]
    labels = {k: np.stack([e[k] for e in examples]) for k in ["aesthetic", "consistency", "jpeg", "neg_jpeg", "labels",
                                                              "weights"] if k in examples[0]}
    tok = lambda t: tokenizer(t, padding="max_length", max_length=tokenizer.model_max_length, return_tensors="np").input_ids
    return {image_field: pixel_values, text_field: tok(captions),
            "idxs": np.stack([e["idx"] for e in examples]), "shuffled_idxs": np.stack([e["shuffled_idx"] for e in examples]),
            "uncond_text": tok([""] * len(examples)), **labels}


class _Loader:
    def __init__(self, dataset, tokenizer, batch_size):
        self.dataset, self.tokenizer, self.batch_size = dataset, tokenizer, batch_size

    def __len__(self):
        return len(self.dataset) // self.batch_size          # drop_last=True

    def __iter__(self):
        for b in range(len(self)):
            yield collate_fn(self.tokenizer, [self.dataset[b * self.batch_size + i] for i in range(self.batch_size)])


def get_bucket_loader(loadpath, tokenizer, batch_size, resolution=None, max_train_samples=None, num_workers=0):
    train_dataset = BucketDataset(utils.ShardReader(loadpath))
    if max_train_samples is not None:
        train_dataset.subsample(max_train_samples)
    train_dataset.with_transform(select_caption)
    train_dataset.shard()
    return train_dataset, _Loader(train_dataset, tokenizer, batch_size)
