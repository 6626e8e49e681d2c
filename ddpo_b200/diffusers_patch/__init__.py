from .scheduling_ddim import DDIMScheduler, DDIMSchedulerState, FlaxDDIMScheduler  # noqa: F401
from .pipeline_stable_diffusion import StableDiffusionPipeline, FlaxStableDiffusionPipeline  # noqa: F401
