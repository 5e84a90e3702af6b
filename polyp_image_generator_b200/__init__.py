"""B200-native DDPM hot path: drop-in for the diffusers objects used by nereaqing/Polyp-Image-Generator.

    from polyp_image_generator_b200 import UNet2DModel, DDPMScheduler, DDPMPipeline, LoraConfig

All tensor arithmetic runs in hand-written sm_100a kernels behind the C-ABI in include/ddpm_b200.h
(libddpm_b200.so, built in-tree by `python -m polyp_image_generator_b200.build`).  There is no CPU fallback.
"""
from .scheduler import (DDIMScheduler, DDIMSchedulerOutput, DDPMScheduler, DDPMSchedulerOutput,  # noqa: F401
                        UniPCMultistepScheduler)
from .pipeline import DDIMPipeline, DDPMPipeline, ImagePipelineOutput, UniPCPipeline, randn_tensor  # noqa: F401


def __getattr__(name):
    # heavier modules are imported lazily so that `import polyp_image_generator_b200` stays cheap
    if name in ("UNet2DModel", "UNet2DOutput"):
        from . import unet
        return getattr(unet, name)
    if name in ("LoraConfig", "lora_state_dict", "recover_lora_modules", "save_lora_weights", "load_lora_weights"):
        from . import lora
        return getattr(lora, name)
    if name in ("mse_loss", "train_step", "train_loop"):
        from . import training
        return getattr(training, name)
    if name in ("PolypGeneratorModel",):
        from . import model
        return getattr(model, name)
    if name in ("FusedAdamW",):
        from . import optim
        return getattr(optim, name)
    if name in ("evaluate", "top_up", "generate_images"):
        from . import sampling
        return getattr(sampling, name)
    if name in ("DistributedDataParallel",):
        from . import ddp
        return getattr(ddp, name)
    raise AttributeError(name)
