#!/usr/bin/env python
"""Per-kernel device time of one denoising step at BASELINE config 5 (AudioLDM-L + r32 LoRA, 750x16 latents, UNet batch 32):
bench.profile_kernels on the c5 pipeline, grouped by kernel and problem shape (top rows)."""
import collections
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import audioldm_with_lora_b200 as b2  # noqa: E402
from audioldm_with_lora_b200 import _lib, synthetic  # noqa: E402

nb, h = 16, 750
cfg = b2.CONFIGS["L"]
unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device="cuda")
unet.load_attn_procs(synthetic.random_lora_state_dict(cfg, 32, fmt="diffusers"))
pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler())
pos, neg = [t.cuda() for t in synthetic.clap_embeddings(nb)]
lat = synthetic.initial_latents(nb, h).cuda()
with torch.no_grad():
    pipe.use_cuda_graph = False
    pipe.denoise(lat, pos, neg, 1, bench.GUIDANCE)
    _lib.PROFILE = []
    pipe.denoise(lat, pos, neg, 1, bench.GUIDANCE)
    torch.cuda.synchronize()
    rec, _lib.PROFILE = _lib.PROFILE, None
rows = collections.defaultdict(lambda: [0, 0.0, 0.0])
for name, e0, e1, info, _ in rec:
    info = info or {}
    key = (name.replace("b200_", ""), info.get("m"), info.get("n"), info.get("k"), info.get("bn"), info.get("taps"), info.get("desc"))
    r = rows[key]; r[0] += 1; r[1] += e0.elapsed_time(e1) * 1e3; r[2] += info.get("flops", 0.0)
tot = sum(r[1] for r in rows.values())
print(f"eager per-launch sum {tot / 1e3:.2f} ms over {sum(r[0] for r in rows.values())} launches")
by = collections.defaultdict(float)
for k, r in rows.items():
    by[k[0]] += r[1]
print({k: round(v / 1e3, 2) for k, v in sorted(by.items(), key=lambda kv: -kv[1])})
for k, r in sorted(rows.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{str(k):92s} {r[0]:3d} x {r[1] / r[0]:8.1f} us = {r[1] / 1e3:7.2f} ms {100 * r[1] / tot:5.1f}%  {r[2] / r[1] / 1e6 if r[1] else 0:6.0f} TFLOP/s")
