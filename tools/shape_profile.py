#!/usr/bin/env python
"""Per-shape device time of one denoising step at the bench workload: every C-ABI call of one eager step is re-issued
8x inside its own CUDA graph and timed with CUDA events (warm L2, PDL active, no Python gaps) -- the same method as
bench.py's roofline leg, grouped by kernel + problem shape."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from audioldm_with_lora_b200 import _lib, synthetic  # noqa: E402

reps = 8
pipe = bench.build_pipeline(torch.device("cuda", 0), 0)
h = 250
pos, neg = [t.cuda() for t in synthetic.clap_embeddings(bench.BATCH)]
lat = synthetic.initial_latents(bench.BATCH, h).cuda()
with torch.no_grad():
    pipe.use_cuda_graph = False
    pipe.denoise(lat, pos, neg, 2, bench.GUIDANCE)
    _lib.PROFILE = []
    pipe.denoise(lat, pos, neg, 1, bench.GUIDANCE)
    torch.cuda.synchronize()
    rec, _lib.PROFILE = _lib.PROFILE, None
    lib = _lib.load()
    side = torch.cuda.Stream()
    rows = {}
    for name, _, _, info, cargs in rec:
        fn = getattr(lib, name)
        g = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            a = list(cargs[:-1]) + [side.cuda_stream]
            with torch.cuda.graph(g, stream=side):
                for _ in range(reps):
                    _lib.check(fn(*a), name)
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side); g.replay(); e1.record(side)
        side.synchronize()
        info = info or {}
        key = (name.replace("b200_", ""), info.get("m"), info.get("n"), info.get("k"), info.get("bn"), info.get("taps"), info.get("desc"))
        d = rows.setdefault(key, {"us": 0.0, "calls": 0, "flops": 0.0})
        d["us"] += e0.elapsed_time(e1) * 1e3 / reps; d["calls"] += 1; d["flops"] += info.get("flops", 0.0)
tot = sum(d["us"] for d in rows.values())
print(f"sum of per-launch times {tot:.1f} us over {sum(d['calls'] for d in rows.values())} launches")
for key, d in sorted(rows.items(), key=lambda kv: -kv[1]["us"]):
    tf = d["flops"] / (d["us"] * 1e-6) / 1e12 if d["flops"] else 0.0
    print(f"{str(key):90s} {d['calls']:3d} x {d['us'] / d['calls']:7.1f} us = {d['us']:8.1f} us  {100 * d['us'] / tot:5.1f}%  {tf:6.0f} TFLOP/s")
