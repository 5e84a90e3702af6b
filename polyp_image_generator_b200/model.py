"""Drop-in for the reference's model wrapper `generator_model/PolypGeneratorModel.py` (from-scratch branch).

    PolypGeneratorModel(device, pretrained=False, add_lora=...)   PolypGeneratorModel.py:13-48
    .get_model()                                                  :50-51
    .add_lora_config(lora_config)                                 :54-58   (unet.add_adapter + trainable-parameter report)
    .unfreeze_layers(layers_to_unfreeze)                          :61-64   (requires_grad = True by name substring)

`pretrained=True` loads the Stable-Diffusion v1.4 VAE / CLIP / UNet2DConditionModel / UniPC from the hub in the
reference; that branch is outside the hot path and raises here.
"""
from __future__ import annotations

from typing import Iterable

from .unet import UNet2DModel


def polyp_unet_config(sample_size: int = 128) -> dict:
    """The UNet2DModel keyword arguments exactly as PolypGeneratorModel.py:26-47 passes them (113 673 219 parameters)."""
    return dict(
        sample_size=sample_size, in_channels=3, out_channels=3, layers_per_block=2,
        block_out_channels=(128, 128, 256, 256, 512, 512),
        down_block_types=("DownBlock2D", "DownBlock2D", "DownBlock2D", "DownBlock2D", "AttnDownBlock2D", "DownBlock2D"),
        up_block_types=("UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"))


def celebahq_unet_config(sample_size: int = 256) -> dict:
    """The google/ddpm-celebahq-256 architecture of BASELINE configs[3]/[4]: the same blocks with
    ONE attention head as wide as the block, downsample_padding 0, [sin, cos] timestep order with freq_shift 1,
    norm_eps 1e-6."""
    cfg = polyp_unet_config(sample_size)
    cfg.update(attention_head_dim=None, downsample_padding=0, flip_sin_to_cos=False, freq_shift=1, norm_eps=1e-6)
    return cfg


class PolypGeneratorModel:
    def __init__(self, device, pretrained: bool = False, add_lora: bool = False, image_size: int = 224):
        """image_size: `TrainingConfig.image_size` of the reference (config_diffusion.py:6 = 224)."""
        self.pretrained = pretrained
        self.add_lora = add_lora
        if pretrained:
            raise NotImplementedError("pretrained=True (Stable Diffusion v1.4 components from the hub) is outside the "
                                      "B200 hot path; construct the from-scratch UNet2DModel with pretrained=False")
        self.unet = UNet2DModel(**polyp_unet_config(image_size))
        if device is not None:
            self.unet.to(device)

    def get_model(self):
        return self.unet

    def add_lora_config(self, lora_config) -> None:
        self.unet.add_adapter(lora_config)
        trainable_params = sum(p.numel() for p in self.unet.parameters() if p.requires_grad)
        total_params = sum(p.numel() for p in self.unet.parameters())
        print(f"Trainable params of unet: {trainable_params} / {total_params} "
              f"({100 * trainable_params / total_params:.2f}%)")

    def unfreeze_layers(self, layers_to_unfreeze: Iterable[str]) -> None:
        layers = list(layers_to_unfreeze)
        for name, param in self.unet.named_parameters():
            if any(x in name for x in layers):
                param.requires_grad = True
