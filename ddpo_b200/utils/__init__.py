"""Host utilities with the names the reference's drivers use (``from ddpo import utils``): ``utils.Parser``,
``utils.Timer``, ``utils.fs``, ``utils.load_unet`` / ``save_unet`` / checkpoints, ``utils.softmax``."""
from . import filesystem as fs  # noqa: F401
from .parser import Parser  # noqa: F401
from .timer import Timer  # noqa: F401
from .serialization import (flat_from_tree, get_latest_epoch, load_finetuned_stable_diffusion,  # noqa: F401
                            load_flax_model, load_unet, n_params, params_tree, restore_checkpoint, save_checkpoint,
                            save_checkpoint_multiprocess, save_unet)
