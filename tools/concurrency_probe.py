#!/usr/bin/env python
"""Do two independent chains of small dependent GEMMs overlap inside one CUDA graph?  (the premise of engine.forward_nhwc(mid_branches=k))"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops, packing  # noqa: E402

dev = "cuda"
m, n, k, depth = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 40
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 0
w = torch.randn(n, k) * k ** -0.5
pw = packing.pack([w], None, 64, 1, k, device=dev)
assert n == k


def chain(x, bufs):
    cur = x
    for i in range(depth):
        ops.conv_gemm(pw, cur, 1, m, 1, bufs[i % 2], max_ctas=cap)
        cur = bufs[i % 2]


def graph_time(nchains):
    xs = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(nchains)]
    bufs = [[torch.empty(m, n, dtype=torch.bfloat16, device=dev) for _ in range(2)] for _ in range(nchains)]
    side = [torch.cuda.Stream() for _ in range(nchains - 1)]
    def body():
        cur = torch.cuda.current_stream()
        streams = [cur] + side
        for s in side:
            s.wait_stream(cur)
        for c, s in enumerate(streams):
            with torch.cuda.stream(s):
                chain(xs[c], bufs[c])
        for s in side:
            cur.wait_stream(s)
    s0 = torch.cuda.Stream()
    s0.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s0):
        body()
    torch.cuda.current_stream().wait_stream(s0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 * 1e3 / depth


for nc in (1, 2, 3, 4):
    print(f"m={m} n={n} k={k} cap={cap}: {nc} chain(s): {graph_time(nc):.2f} us per chain-link ({depth} links)", flush=True)
