from .schedulers import (DDIMScheduler, DDIMInverseScheduler, DDPMInverseScheduler, DiffusionInverseScheduler,  # noqa: F401
                         FrozenConfig)
