"""Host utilities with the names the reference's drivers use (``from ddpo import utils``): ``utils.Parser``,
``utils.Timer``, ``utils.fs``, ``utils.load_unet`` / ``save_unet`` / checkpoints, ``utils.softmax``."""
from . import filesystem as fs  # noqa: F401
from .parser import Parser  # noqa: F401
from .timer import Timer  # noqa: F401
from .serialization import (flat_from_tree, get_latest_epoch, load_finetuned_stable_diffusion,  # noqa: F401
                            load_flax_model, load_unet, n_params, params_tree, restore_checkpoint, save_checkpoint,
                            save_checkpoint_multiprocess, save_unet)
from .logger import Masker, Percentile, StreamingAverage, StreamingPercentile, Threshold, make_masker  # noqa: F401
from .shards import (ShardReader, ShardWriter, decode_generic, decode_jpeg, encode_generic, encode_jpeg,  # noqa: F401
                     softmax_ref)


def softmax(x, temperature=1.0):
    """``utils.softmax`` (reference ``ddpo/utils/array.py:44-56``): softmax over ALL devices' entries -- with one
    process per GPU the per-worker array is the whole local batch; the cross-rank max / sum use torch.distributed."""
    import numpy as np
    import torch
    from ..training import distributed
    x = np.asarray(x, np.float64) * temperature
    if distributed.world_size() == 1:
        return softmax_ref(x.reshape(-1), 1.0).reshape(x.shape)
    m = torch.tensor([x.max()], dtype=torch.float64)
    torch.distributed.all_reduce(m, op=torch.distributed.ReduceOp.MAX)
    e = np.exp(x - m.item())
    s = torch.tensor([e.sum()], dtype=torch.float64)
    torch.distributed.all_reduce(s)
    return e / s.item()
