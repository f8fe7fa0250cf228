#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last full denoising step
(kernels between the last two sampler_step launches), grouped by kernel and, with a layers.json from
tools/layer_profile.py, by layer shape.   python tools/ncu_step.py launches.csv [layers.json] [top]"""
import collections
import csv
import json
import sys


def main():
    path = sys.argv[1]
    layers = sys.argv[2] if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = [(x["Kernel Name"], float(x["Metric Value"]), x["Grid Size"]) for x in csv.DictReader(lines)]
    idx = [i for i, (n, _, _) in enumerate(rows) if "sampler_step" in n]
    # the last pair of consecutive sampler launches with only this library's kernels in between
    mine = ("conv_gemm", "attention_kernel", "groupnorm", "gn_", "layernorm", "time_class_embed", "upsample_nearest", "sampler_step", "splitk_reduce")
    step = None
    for a, b in reversed(list(zip(idx[:-1], idx[1:]))):
        cand = rows[a + 1: b + 1]
        if all(any(m in n for m in mine) for n, _, _ in cand):
            step = cand
            break
    if step is None:
        raise SystemExit("no clean denoising step found in the launch list")
    total = sum(t for _, t, _ in step) / 1e3
    print(f"kernels in one denoising step: {len(step)}; sum of durations {total:.1f} us")
    by = collections.defaultdict(lambda: [0, 0.0])
    for n, t, g in step:
        k = n.split("(")[0].replace("void ", "").replace("b200::", "")[:48]
        by[k][0] += 1; by[k][1] += t
    for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:<50} {c:>4} launches {t / 1e3:9.1f} us  share {t / 1e3 / total * 100:5.1f}%  avg {t / c / 1e3:7.2f} us")
    if not layers:
        return
    L = json.load(open(layers))["rows"]
    # a split-K layer is one call but two kernels (GEMM + reduce): fold the reduce into the preceding kernel
    folded = []
    for n, t, g in step:
        if "splitk_reduce" in n and folded:
            folded[-1] = (folded[-1][0], folded[-1][1] + t, folded[-1][2])
        else:
            folded.append((n, t, g))
    step = folded
    if len(L) != len(step):
        print(f"(layers.json has {len(L)} calls, step has {len(step)} kernels: no per-layer table)")
        return
    by = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r, (_, t, _) in zip(L, step):
        k = (r["name"], r.get("m"), r.get("n"), r.get("k"), r.get("bn"), r.get("taps"), r.get("desc"))
        by[k][0] += 1; by[k][1] += t / 1e3; by[k][2] += r.get("flops", 0)
    for k, (c, t, f) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {str(k):<74} {c:>3} x {t / c:7.1f} us = {t:8.1f} us  {f / t / 1e6 if f else 0:6.0f} TFLOP/s")


if __name__ == "__main__":
    main()
