"""ctypes binding of libb200ldm.so (the C-ABI declared in include/b200ldm.h).

There is NO fallback: if the shared library is missing or a kernel call fails, this raises.
PyTorch is only the owner of device memory and streams; every compute call below lands in a
hand-written sm_100a kernel.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_long, c_void_p
from pathlib import Path
from typing import Optional

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("B200LDM_LIB", _PKG / "libb200ldm.so"))

_lib: Optional[ctypes.CDLL] = None
launch_count = 0          # kernels launched through this binding (bench.py's gpu_launches)


class B200Error(RuntimeError):
    pass


_PROTOS = {
    "b200_version": (c_int, []),
    "b200_last_error": (c_char_p, []),
    "b200_debug_timeline": (c_int, [c_void_p, c_int]),
    "b200_set_sm_budget": (c_int, [c_int]),
    "b200_tmap_cache_hits": (c_long, []),
    "b200_conv_gemm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                               c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "b200_conv_gemm_gnstat": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                      c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_gn_stat_slabs": (c_int, [c_int, c_int, c_int]),
    "b200_gemm_nt": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_conv1d": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                            c_int, c_float, c_void_p, c_int, c_long, c_int, c_float, c_int, c_int, c_int, c_void_p]),
    "b200_conv3x3_s2_pad01": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                      c_int, c_void_p]),
    "b200_lrelu_mean3": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_float, c_float, c_void_p, c_void_p]),
    "b200_f32_to_bf16": (c_int, [c_void_p, c_long, c_void_p, c_void_p]),
    "b200_groupnorm_apply": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p]),
    "b200_linear_lora": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                 c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "b200_linear_ln": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p,
                               c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p]),
    "b200_linear_stats": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int,
                                  c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "b200_groupnorm_silu": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float,
                                    c_int, c_void_p, c_void_p]),
    "b200_softmax_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_long, c_void_p, c_long, c_float, c_void_p]),
    "b200_layernorm": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "b200_embed_layernorm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_float, c_void_p, c_void_p]),
    "b200_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "b200_time_class_embed": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200_pack_nchw_to_nhwc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_unpack_nhwc_to_nchw": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_upsample_nearest": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_sampler_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int,
                                  c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_add_noise": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_adamw_flat": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_float, c_float, c_float, c_float,
                                c_float, c_int, c_float, c_void_p]),
    "b200_mse_partial": (c_int, [c_void_p, c_void_p, c_long, c_void_p, c_void_p]),
    "b200_adamw_flat_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_float, c_float, c_float,
                                    c_float, c_void_p]),
    # ---- fine-tuning step (forward variants that keep what the backward needs, and the backward kernels)
    "b200_attention_lse": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "b200_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_float, c_void_p]),
    "b200_groupnorm_silu_stats": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                          c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "b200_groupnorm_silu_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                        c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b200_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_void_p]),
    "b200_geglu_fwd": (c_int, [c_void_p, c_long, c_int, c_void_p, c_void_p]),
    "b200_geglu_bwd": (c_int, [c_void_p, c_void_p, c_long, c_int, c_void_p, c_void_p]),
    "b200_lora_wgrad": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "b200_zero_insert": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_upsample_nearest_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_add_bf16": (c_int, [c_void_p, c_void_p, c_long, c_void_p]),
    "b200_mse_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "b200_lora_refresh": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
}
EXPORTS = tuple(_PROTOS)


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise B200Error(f"{LIB_PATH} not found: build it with `python -m audioldm_with_lora_b200.build` "
                            "(there is no CPU / PyTorch fallback)")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().b200_last_error().decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise B200Error(f"{what} failed (code {rc}): {last_error()}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise B200Error("b200 kernels need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


PROFILE = None            # set to a list to record (name, start_event, end_event, info, args) per call


def call(name: str, *args, info=None) -> None:
    global launch_count
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(load(), name)(*args)
    check(rc, name)
    if PROFILE is not None:
        e1.record()
        PROFILE.append((name, e0, e1, info, args))
    launch_count += 1
