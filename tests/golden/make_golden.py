#!/usr/bin/env python
"""Generate tests/golden/*.npz from the fp32 CPU oracle (run here, committed; the GPU box never needs
/root/reference or this script).

The reference holds NO golden vectors or tests for this path and its arithmetic lives in un-vendored
third-party packages (diffusers 0.32.2 / peft 0.13.2, not installable offline), so these fixtures pin
the oracle restatement itself ("parity unpinned") -- they guard against silent drift of
the oracle and give the GPU tests a reference that does not need the oracle at run time.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from audioldm_with_lora_b200 import synthetic                      # noqa: E402
from audioldm_with_lora_b200.arch import CONFIGS                   # noqa: E402
from audioldm_with_lora_b200.lora import parse_lora_state_dict     # noqa: E402
from oracle import pipeline_ref, unet_ref                          # noqa: E402
from oracle.ddim_ref import DDIMRef, PNDMRef                       # noqa: E402

OUT = Path(__file__).resolve().parent
TF_STEPS = (0, 1, 2, 5, 10, 25, 50, 100, 150, 199)


def main():
    torch.set_num_threads(8)
    cfg = CONFIGS["S"]
    sd = synthetic.random_unet_state_dict(cfg, seed=0)
    ad = parse_lora_state_dict(synthetic.random_lora_state_dict(cfg, 8, fmt="peft"))
    lora = unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()})
    # 1. one UNet call: S arch, r8 LoRA, batch 2, latent 24x16 (odd sizes down the pyramid: 24,12,6,3)
    x = synthetic.initial_latents(2, 24)
    pos, neg = synthetic.clap_embeddings(2)
    with torch.no_grad():
        eps = unet_ref.unet_forward(sd, unet_ref.ARCH_S, x, 501, pos, lora=lora)
        eps_nolora = unet_ref.unet_forward(sd, unet_ref.ARCH_S, x, 501, pos)
        # per-sample timesteps as in the fine-tuning step (train_audioldm_lora.py:499-504,539)
        eps_t = unet_ref.unet_forward(sd, unet_ref.ARCH_S, x, torch.tensor([17, 903]), pos, lora=lora)
    np.savez_compressed(OUT / "unet_s_r8_b2_h24.npz", eps=eps.numpy(), eps_nolora=eps_nolora.numpy(),
                        eps_pert=eps_t.numpy())
    # 2. a 4-step CFG DDIM trajectory, 1 prompt, latent 25x16 (5 s clip / 5), guidance 2.5
    x1 = synthetic.initial_latents(1, 25)
    p1, n1 = synthetic.clap_embeddings(1)
    trace = []
    with torch.no_grad():
        pipeline_ref.denoise_loop(sd, unet_ref.ARCH_S, p1, n1, x1.clone(), 4, 2.5, lora=lora, trace=trace)
    np.savez_compressed(OUT / "ddim_s_r8_b1_h25_4steps.npz", latents=torch.stack(trace).numpy())
    # 2b. BASELINE config c2's own latent size: 200 CFG DDIM steps, 1 prompt, latent 250x16 (10 s clip), guidance 2.5.
    #     Kept: the latent BEFORE step k and AFTER it for k in TF_STEPS (teacher-forced per-step check on the GPU:
    #     feed `before`, run one step, compare with `after`) and the final latent (free-running drift, log-mel L1).
    x2 = synthetic.initial_latents(1, 250)
    trace2 = [x2.clone()]                  # DDIM init_noise_sigma = 1: index k = latent before step k
    with torch.no_grad():
        pipeline_ref.denoise_loop(sd, unet_ref.ARCH_S, p1, n1, x2.clone(), 200, 2.5, lora=lora, trace=trace2)
    keep = sorted({k for s_ in TF_STEPS for k in (s_, s_ + 1)} | {200})
    np.savez_compressed(OUT / "ddim_s_r8_b1_h250_200steps.npz", index=np.array(keep, dtype=np.int32),
                        latents=torch.stack([trace2[k] for k in keep]).numpy())
    # 3. scheduler known answers
    s = DDIMRef()
    np.savez_compressed(OUT / "ddim_schedule.npz", alphas_cumprod=s.alphas_cumprod.numpy(),
                        t200=s.set_timesteps(200).numpy(), t50=s.set_timesteps(50).numpy(), t10=s.set_timesteps(10).numpy(),
                        plms10=PNDMRef().set_timesteps(10).numpy())
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))


if __name__ == "__main__":
    main()
