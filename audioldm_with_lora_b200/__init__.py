"""B200-native hot path of AudioLDM-with-LoRA: the LoRA-adapted UNet denoising loop behind the
diffusers attention-processor / LoRA-loader API and the AudioLDMPipeline call signature.

Compute runs in hand-written sm_100a kernels (csrc/, C-ABI in include/b200ldm.h); PyTorch owns
device memory and streams only.  There is no CPU or PyTorch fallback for the hot path.
"""
from . import clap, vae, vocoder  # noqa: F401  (b2.clap.from_torch_text_encoder, b2.vae.from_torch_encoder, ...)
from .arch import AUDIOLDM_L, AUDIOLDM_S, CONFIGS, UNetConfig
from .clap import B200ClapTextEncoder
from .lora import (LoraConfig, convert_state_dict_to_diffusers, load_lora_checkpoint, merge_lora_into_state_dict,
                   parse_lora_state_dict, save_lora_checkpoint, to_peft_state_dict)
from .model import (Attention, B200AttnProcessor, LoraLinear, UNet2DConditionModel, UNet2DConditionOutput,
                    get_peft_model, get_peft_model_state_dict)
from .pipeline import AudioLDMPipeline, AudioPipelineOutput
from .scheduler import DDIMScheduler, PNDMScheduler
from .vae import B200VaeDecoder, B200VaeEncoder
from .vocoder import B200HifiGan

__all__ = [
    "AUDIOLDM_L", "AUDIOLDM_S", "CONFIGS", "UNetConfig", "LoraConfig", "convert_state_dict_to_diffusers",
    "parse_lora_state_dict", "to_peft_state_dict", "save_lora_checkpoint", "load_lora_checkpoint",
    "merge_lora_into_state_dict", "Attention", "B200AttnProcessor", "LoraLinear",
    "UNet2DConditionModel", "UNet2DConditionOutput", "get_peft_model", "get_peft_model_state_dict",
    "AudioLDMPipeline", "AudioPipelineOutput", "DDIMScheduler", "PNDMScheduler", "B200VaeDecoder", "B200VaeEncoder", "B200HifiGan",
    "B200ClapTextEncoder",
]
