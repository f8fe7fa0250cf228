#!/usr/bin/env python
"""Which kernel type breaks the overlap of independent chains inside one CUDA graph?  Chains of `kind` links."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops, packing  # noqa: E402

dev = "cuda"
nb, hw, c, depth = 4, 64, 640, 24
m = nb * hw
w = torch.randn(c, c) * c ** -0.5
pw = packing.pack([w], None, 64, 1, c, device=dev)
pw3 = packing.pack([torch.randn(3 * c, c) * c ** -0.5], None, 128, 1, c, device=dev)
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)


def link(kind, x, y, qkv):
    if kind == "gemm":
        ops.conv_gemm(pw, x, 1, m, 1, y)
    elif kind == "ln":
        ops.layernorm(x, m, c, gamma, beta, 1e-5, y)
    elif kind == "gn":
        ops.groupnorm_silu(x, c, None, 0, nb, hw, gamma, beta, 1e-5, True, y)
    elif kind == "attn":
        ops.conv_gemm(pw3, x, 1, m, 1, qkv)
        ops.attention(qkv, y, nb, hw, 8, c // 8)
    elif kind == "mix":
        ops.layernorm(x, m, c, gamma, beta, 1e-5, y)
        ops.conv_gemm(pw3, y, 1, m, 1, qkv)
        ops.attention(qkv, x, nb, hw, 8, c // 8)
        ops.conv_gemm(pw, x, 1, m, 1, y)
        ops.groupnorm_silu(y, c, None, 0, nb, hw, gamma, beta, 1e-5, True, x)
        ops.conv_gemm(pw, x, 1, m, 1, y)


def graph_time(kind, nchains):
    xs = [torch.randn(m, c, device=dev).to(torch.bfloat16) for _ in range(nchains)]
    bufs = [[torch.empty(m, c, dtype=torch.bfloat16, device=dev) for _ in range(2)] for _ in range(nchains)]
    qkvs = [torch.empty(m, 3 * c, dtype=torch.bfloat16, device=dev) for _ in range(nchains)]
    side = [torch.cuda.Stream() for _ in range(nchains - 1)]

    def body():
        cur = torch.cuda.current_stream()
        streams = [cur] + side
        for s in side:
            s.wait_stream(cur)
        for ci, s in enumerate(streams):
            with torch.cuda.stream(s):
                a, b = xs[ci], bufs[ci][0]
                for i in range(depth):
                    link(kind, a, b, qkvs[ci])
                    a, b = b, (bufs[ci][1] if b is bufs[ci][0] else bufs[ci][0])
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


import os
if os.environ.get("PROBE_BUDGET"):
    ops.set_sm_budget(int(os.environ["PROBE_BUDGET"]))
for kind in sys.argv[1:] or ["gemm", "ln", "gn", "attn", "mix"]:
    print(kind, "  ".join(f"{nc} chains: {graph_time(kind, nc):7.2f} us/link" for nc in (1, 2, 4)), flush=True)
