from .schedulers import DDIMScheduler, DDIMInverseScheduler, DiffusionInverseScheduler, FrozenConfig  # noqa: F401
