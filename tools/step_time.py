#!/usr/bin/env python
"""Graph-replayed denoising step time at the bench workload (median of 40 replays after 10)."""
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import audioldm_with_lora_b200 as b2  # noqa: E402
from audioldm_with_lora_b200 import synthetic  # noqa: E402

cfg = b2.CONFIGS["S"]
unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device="cuda")
unet.load_state_dict(synthetic.random_lora_state_dict(cfg, 8, fmt="peft"), strict=False)
lat = synthetic.initial_latents(8, 250).cuda()
pos, neg = [t.cuda() for t in synthetic.clap_embeddings(8)]
ref = None
for br in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), branches=br)
    out = pipe.denoise(lat, pos, neg, 2, 2.5)
    if ref is None:
        ref = out
    dev = ((out - ref).norm() / ref.norm()).item()
    st = next(iter(pipe._loops.values()))
    ms = []
    for _ in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.step.zero_()
        e0.record(); st.graph.replay(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(f"branches {br}: graph step median {statistics.median(ms[10:]):.3f} ms  min {min(ms):.3f} ms   "
          f"rel-L2 vs first config {dev:.2e}", flush=True)
