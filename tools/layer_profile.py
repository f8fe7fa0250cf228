#!/usr/bin/env python
"""Per-launch profile of one denoising step at the bench workload (config 2: UNet batch 16 @ 250x16).

    python tools/layer_profile.py [out.json] [reps=5] [sweep=1]

Eager (no CUDA graph) step with CUDA events around every C-ABI call, `reps` times, median per call;
then an optional block_n / CTA-count sweep of the heaviest GEMM shapes with back-to-back launches
(kernel-only time).  Writes a JSON summary; prints a table sorted by time.
"""
from __future__ import annotations

import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import audioldm_with_lora_b200 as b2  # noqa: E402
from audioldm_with_lora_b200 import _lib, ops, packing, synthetic  # noqa: E402


def time_fn(fn, reps=20, warm=3, graph=True):
    """ms per call.  graph=True: `reps` launches captured in one CUDA graph and replayed (no Python / ctypes launch
    overhead between kernels; back-to-back launches overlap through PDL exactly as inside the denoising step)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(reps):
                    fn()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and "=" not in sys.argv[1] else "gpurun_out/layers.json"
    kw = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    reps, sweep = int(kw.get("reps", 5)), int(kw.get("sweep", 1))
    batch, h = int(kw.get("batch", 8)), int(kw.get("h", 250))
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device="cuda")
    unet.load_state_dict(synthetic.random_lora_state_dict(cfg, 8, fmt="peft"), strict=False)
    pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), use_cuda_graph=False)
    lat = synthetic.initial_latents(batch, h).cuda()
    pos, neg = [t.cuda() for t in synthetic.clap_embeddings(batch)]
    pipe.denoise(lat, pos, neg, 2, 2.5)
    runs = []
    for _ in range(reps):
        _lib.PROFILE = []
        pipe.denoise(lat, pos, neg, 1, 2.5)
        torch.cuda.synchronize()
        rec, _lib.PROFILE = _lib.PROFILE, None
        runs.append([(n, e0.elapsed_time(e1), info) for n, e0, e1, info, _ in rec])
    rows = []
    for i, (name, _, info) in enumerate(runs[0]):
        ms = statistics.median(r[i][1] for r in runs)
        row = {"i": i, "name": name.replace("b200_", ""), "ms": round(ms, 5)}
        row.update(info or {})
        if info and "flops" in info:
            row["tflops"] = round(info["flops"] / ms / 1e9, 1)
        rows.append(row)
    total = sum(r["ms"] for r in rows)
    by = {}
    for r in rows:
        key = (r["name"], r.get("m"), r.get("n"), r.get("k"), r.get("bn"), r.get("taps"), r.get("desc"))
        d = by.setdefault(key, {"calls": 0, "ms": 0.0, "flops": 0.0})
        d["calls"] += 1; d["ms"] += r["ms"]; d["flops"] += r.get("flops", 0.0)
    print(f"total event-timed {total:.3f} ms over {len(rows)} calls")
    table = []
    for key, d in sorted(by.items(), key=lambda kv: -kv[1]["ms"]):
        tf = d["flops"] / d["ms"] / 1e9 if d["flops"] else 0.0
        table.append({"key": [k for k in key], **{k: round(v, 4) for k, v in d.items()}, "tflops": round(tf, 1)})
        print(f"{key[0]:<18} m={key[1]} n={key[2]} k={key[3]} bn={key[4]} taps={key[5]} {key[6] or ''}: "
              f"{d['calls']:>3} calls {d['ms']:.4f} ms ({d['ms'] / d['calls'] * 1e3:.1f} us each) {tf:.0f} TFLOP/s")
    result = {"total_ms": total, "calls": len(rows), "groups": table, "rows": rows}

    if sweep:
        # graph replay time of the whole step for comparison
        pipe.use_cuda_graph = True
        pipe.denoise(lat, pos, neg, 2, 2.5)
        st = next(iter(pipe._loops.values()))
        result["graph_step_ms"] = time_fn(lambda: st.graph.replay(), reps=20, graph=False)
        print("graph step ms", result["graph_step_ms"])
        sw = []
        shapes = [  # (label, nb, h, w, cin, cout, taps)
            ("conv L0 128->128", 16, 250, 16, 128, 128, 9), ("conv L0 256->128", 16, 250, 16, 256, 128, 9),
            ("conv L1 256->256", 16, 125, 8, 256, 256, 9), ("conv L2 384->384", 16, 63, 4, 384, 384, 9),
            ("conv L3 640->640", 16, 32, 2, 640, 640, 9), ("conv L3 1280->640", 16, 32, 2, 1280, 640, 9),
            ("lin L1 qkv 256->768", 1, 16000, 1, 256, 768, 1), ("lin L1 ff1 256->2048", 1, 16000, 1, 256, 2048, 1),
            ("lin L1 ff2 1024->256", 1, 16000, 1, 1024, 256, 1), ("lin L1 256->256", 1, 16000, 1, 256, 256, 1),
            ("lin L3 qkv 640->1920", 1, 1024, 1, 640, 1920, 1), ("lin L3 640->640", 1, 1024, 1, 640, 640, 1),
            ("lin L2 384->384", 1, 4032, 1, 384, 384, 1),
        ]
        g = torch.Generator().manual_seed(0)
        for label, nb, hh, ww, ci, co, taps in shapes:
            x = torch.randn(nb * hh * ww, ci, generator=g).to("cuda", torch.bfloat16)
            wt = torch.randn(co, taps * ci, generator=g) * (taps * ci) ** -0.5
            out = torch.empty(nb * hh * ww, co, dtype=torch.bfloat16, device="cuda")
            flops = 2.0 * nb * hh * ww * co * taps * ci
            for bn in (64, 128, 192, 256):
                if bn > co and bn != 32:
                    continue
                pw = packing.pack([wt], torch.zeros(co), bn, taps, ci, device="cuda")
                ms = time_fn(lambda: ops.conv_gemm(pw, x, nb, hh, ww, out))
                mt = ops.num_m_tiles(nb, hh, ww)
                sw.append({"shape": label, "bn": bn, "tiles": mt * (pw.n_pad // bn), "us": round(ms * 1e3, 2),
                           "tflops": round(flops / ms / 1e9, 1)})
                print(sw[-1])
        result["sweep"] = sw
        # attention + norms, kernel-only
        misc = []
        for (b, s, d) in [(16, 1000, 32), (16, 252, 48), (16, 64, 80)]:
            qkv = torch.randn(b, s, 3 * 8 * d, generator=g).to("cuda", torch.bfloat16)
            o = torch.empty(b, s, 8 * d, dtype=torch.bfloat16, device="cuda")
            ms = time_fn(lambda: ops.attention(qkv, o, b, s, 8, d))
            misc.append({"op": f"attention b{b} s{s} d{d}", "us": round(ms * 1e3, 2),
                         "tflops": round(4.0 * b * 8 * s * s * d / ms / 1e9, 1)})
        for (nb, hw, c) in [(16, 4000, 128), (16, 4000, 256), (16, 1000, 256), (16, 1000, 512), (16, 252, 384), (16, 64, 640), (16, 64, 1280)]:
            x = torch.randn(nb, hw, c, generator=g).to("cuda", torch.bfloat16)
            y = torch.empty_like(x)
            gm, bt = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
            ms = time_fn(lambda: ops.groupnorm_silu(x, c, None, 0, nb, hw, gm, bt, 1e-5, True, y))
            misc.append({"op": f"groupnorm nb{nb} hw{hw} c{c}", "us": round(ms * 1e3, 2),
                         "gbs": round(nb * hw * c * 2 * 2 / ms / 1e6, 1)})
        for (m, c) in [(16000, 256), (4032, 384), (1024, 640)]:
            x = torch.randn(m, c, generator=g).to("cuda", torch.bfloat16)
            y = torch.empty_like(x)
            gm, bt = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
            ms = time_fn(lambda: ops.layernorm(x, m, c, gm, bt, 1e-5, y))
            misc.append({"op": f"layernorm m{m} c{c}", "us": round(ms * 1e3, 2), "gbs": round(m * c * 2 * 2 / ms / 1e6, 1)})
        for r in misc:
            print(r)
        result["misc"] = misc
    Path(out_path).parent.mkdir(exist_ok=True, parents=True)
    Path(out_path).write_text(json.dumps(result, indent=1))


if __name__ == "__main__":
    main()
