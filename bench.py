#!/usr/bin/env python
"""bench.py -- generated audio-sec/sec of the LoRA-adapted AudioLDM sampling path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): AudioLDM-S architecture (random-init, seed 0) + rank-8 LoRA on
attn q/k/v/out, batch 8 prompts per GPU, 200 DDIM steps, CFG 2.5, 10 s clips, bf16 kernels with fp32
accumulation, synthetic L2-normalised CLAP embeddings.  One bench "step" = one batch of 8 clips
through the whole path (200 x {CFG-doubled UNet, guidance, DDIM update} + VAE decode + vocoder).

  value  : B*10 s*K / device time, inputs already resident in HBM (denoise loop + the torch VAE/vocoder tail, both replayed from CUDA graphs)
  e2e    : the same through the public call `AudioLDMPipeline.__call__(prompt_embeds=<host tensors>, ...)`
           returning host numpy audio (H2D of embeddings/latents and D2H of waveforms inside the timing)
  roofline: the implicit-GEMM kernel (all conv / linear layers, ~89 % of the step's FLOPs) timed per
           launch with CUDA events on the launching stream; algorithmic FLOPs / time vs measured bf16 peak
  cpu_baseline: the torch-eager fp32 oracle (restating the reference's diffusers path) on the host cores,
           bounded sample, extrapolated to the same 200-step / 10 s clip.
--impl reference prints that CPU arm as its own line (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CLIP_S = 10.0
STEPS_DDIM = 200
BATCH = 8
GUIDANCE = 2.5
RANK_LORA = 8
METRIC = "generated_audio_sec_per_sec"
UNIT = "audio-s/s"
CONFIG = {"workload": "AudioLDM-S (audioldm-s-full-v2 arch, random-init) + rank-8 LoRA q/k/v/out, batch 8/GPU, "
                      "200 DDIM steps, CFG 2.5, 10 s clips", "batch_per_gpu": BATCH, "ddim_steps": STEPS_DDIM,
          "guidance_scale": GUIDANCE, "clip_s": CLIP_S, "lora_rank": RANK_LORA, "parallelism": "prompt-sharded, no collectives",
          "l2_policy": "per-step working set (370 MB bf16 weights + activations) exceeds the 126 MB L2; no flush needed"}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1412.4), d.get("bf16_tflops", 1661.0), d.get("hbm_gbs", 6533.5), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(steps: int, warmup: int) -> dict:
    """The reference's CPU path (torch-eager fp32 oracle of diffusers/peft semantics) on the host cores."""
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.arch import CONFIGS
    from audioldm_with_lora_b200.lora import parse_lora_state_dict
    from oracle import pipeline_ref, unet_ref, vae_ref
    from oracle.ddim_ref import DDIMRef
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = CONFIGS["S"]
    sd = synthetic.random_unet_state_dict(cfg, seed=0)
    ad = parse_lora_state_dict(synthetic.random_lora_state_dict(cfg, RANK_LORA, fmt="peft"))
    lora = unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()})
    h = pipeline_ref.latent_height(CLIP_S)
    lat = synthetic.initial_latents(1, h)
    pos, neg = synthetic.clap_embeddings(1)
    emb = torch.cat([neg, pos])
    sched = DDIMRef(); ts = sched.set_timesteps(STEPS_DDIM)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            t = ts[i]
            e = unet_ref.unet_forward(sd, unet_ref.ARCH_S, torch.cat([lat] * 2), t, emb, lora=lora)
            eu, et = e.chunk(2)
            lat = sched.step(eu + GUIDANCE * (et - eu), int(t), lat)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        # tail once: VAE decode + vocoder for the one clip
        from audioldm_with_lora_b200.tail import build_vocoder
        vae_sd = synthetic.random_state_dict_from_shapes(vae_ref.vae_decoder_param_shapes(), seed=7)
        voc = build_vocoder(0)
        t0 = time.perf_counter()
        pipeline_ref.decode_tail(vae_sd, voc, lat, CLIP_S)
        tail = time.perf_counter() - t0
    step_s = sum(times) / len(times)
    clip_time = STEPS_DDIM * step_s + tail
    value = CLIP_S / clip_time
    sample = (f"1 prompt, 10 s clip: {steps} CFG denoising steps (UNet batch 2, fp32, unmerged LoRA) timed after {warmup} "
              f"warm-up, mean {step_s * 1e3:.0f} ms/step, + VAE decode and vocoder once ({tail:.1f} s); extrapolated to "
              f"{STEPS_DDIM} steps")
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_denoise_step": step_s * 1e3,
            "tail_s": tail}


def gpu_eager_arm(device, reps: int = 5) -> dict:
    """Same-box GPU baseline: the torch-eager oracle -- the ATen ops diffusers would issue (cuDNN convs,
    cuBLAS linears, SDPA, native GroupNorm / LayerNorm, ~5 launches per unmerged LoRA linear) -- on this GPU in bf16, one
    CFG-doubled UNet call at the bench shape.  A reported baseline like cpu_baseline, not the product path."""
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.arch import CONFIGS
    from audioldm_with_lora_b200.lora import parse_lora_state_dict
    from oracle import unet_ref
    cfg = CONFIGS["S"]
    dt = torch.bfloat16
    sd = {k: v.to(device, dt) for k, v in synthetic.random_unet_state_dict(cfg, seed=0).items()}
    ad = parse_lora_state_dict(synthetic.random_lora_state_dict(cfg, RANK_LORA, fmt="peft"))
    lora = unet_ref.LoraSet({k: (e.A.to(device, dt), e.B.to(device, dt), e.alpha) for k, e in ad.items()})
    h = int(CLIP_S / 0.01) // 4
    x = torch.randn(2 * BATCH, 8, h, 16, device=device, dtype=dt)
    emb = torch.nn.functional.normalize(torch.randn(2 * BATCH, 512, device=device), dim=-1).to(dt)
    with torch.no_grad():
        for _ in range(3):
            unet_ref.unet_forward(sd, unet_ref.ARCH_S, x, 500, emb, lora=lora)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            unet_ref.unet_forward(sd, unet_ref.ARCH_S, x, 500, emb, lora=lora)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"unet_step_ms": ms, "kind": "oracle port on this GPU: torch eager bf16 (cuDNN / cuBLAS / SDPA), unmerged LoRA, UNet batch 16",
            "implied_audio_s_per_s_loop_only": BATCH * CLIP_S / (STEPS_DDIM * ms / 1e3)}


# ------------------------------------------------------------------------------------------ B200 arm
def build_pipeline(device, rank_seed_base: int):
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic, tail
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=device)
    unet.load_state_dict(synthetic.random_lora_state_dict(cfg, RANK_LORA, fmt="peft"), strict=False)
    pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), vae=tail.random_vae_decoder(7), vocoder=tail.build_vocoder(0))
    return pipe


def profile_kernels(pipe, lat, pos, neg, reps: int = 8):
    """Per-kernel device time of one denoising step.  One eager step records every C-ABI call with its arguments;
    each call is then re-issued `reps` times inside its own CUDA graph and timed with CUDA events on the launching
    stream (device time per launch without Python / ctypes launch gaps, programmatic dependent launch active as in
    the real step graph, operands L2-warm)."""
    from audioldm_with_lora_b200 import _lib
    pipe.use_cuda_graph = False
    pipe.denoise(lat, pos, neg, 2, GUIDANCE)            # warm (weights packed, attributes set)
    _lib.PROFILE = []
    n0 = _lib.launch_count
    pipe.denoise(lat, pos, neg, 1, GUIDANCE)
    torch.cuda.synchronize()
    launches = _lib.launch_count - n0
    rec, _lib.PROFILE = _lib.PROFILE, None
    pipe.use_cuda_graph = True
    lib = _lib.load()
    by = {}
    side = torch.cuda.Stream()
    for name, _, _, info, cargs in rec:
        fn = getattr(lib, name)
        # the recorded stream argument (last) is replaced by the capture stream
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
        d = by.setdefault(name, {"ms": 0.0, "calls": 0, "flops": 0.0})
        d["ms"] += e0.elapsed_time(e1) / reps; d["calls"] += 1
        d["flops"] += (info or {}).get("flops", 0.0)
    torch.cuda.current_stream().wait_stream(side)
    return by, launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ddim-steps", type=int, default=STEPS_DDIM,
                    help="denoising steps per clip (default 200 = the BASELINE config; smaller only for ncu launch lists)")
    args = ap.parse_args()
    if args.ddim_steps != STEPS_DDIM:
        globals()["STEPS_DDIM"] = args.ddim_steps
        CONFIG["ddim_steps"] = args.ddim_steps
        CONFIG["workload"] += f" [NOT THE BASELINE CONFIG: {args.ddim_steps} DDIM steps]"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_arm(max(args.steps, 1), args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_denoise_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the hot path has no CPU fallback")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    from audioldm_with_lora_b200 import _lib, synthetic

    pipe = build_pipeline(device, rank)
    h = int(CLIP_S / 0.01) // 4
    # prompts are sharded by global index: rank r owns prompts [r*B, (r+1)*B)  (no data-path collective)
    from audioldm_with_lora_b200.sharding import shard_prompts
    pos_h, neg_h = synthetic.clap_embeddings(BATCH * world)
    pos_h, neg_h, owned = shard_prompts(pos_h, neg_h, rank, world)
    pos_h, neg_h = pos_h.contiguous().pin_memory(), neg_h.contiguous().pin_memory()
    lat_h = synthetic.initial_latents(len(owned), h, first_index=owned[0]).pin_memory()
    pos_d, neg_d, lat_d = pos_h.to(device), neg_h.to(device), lat_h.to(device)

    def resident_step():
        lat = pipe.denoise(lat_d, pos_d, neg_d, STEPS_DDIM, GUIDANCE)
        return pipe.latents_to_waveform(lat)

    def e2e_step():
        return pipe(prompt_embeds=pos_h, negative_prompt_embeds=neg_h, latents=lat_h, audio_length_in_s=CLIP_S,
                    num_inference_steps=STEPS_DDIM, guidance_scale=GUIDANCE).audios

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    with torch.no_grad():
        for _ in range(args.warmup):
            resident_step()
        sampler = ClockSampler(local_rank)
        sampler.start()
        ms_res = timed(resident_step, args.steps)
        sampler.stop_flag = True
        for _ in range(min(args.warmup, 1)):
            e2e_step()
        # e2e: wall clock around the public call (it ends with a D2H copy, so the host clock is exact)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            audio = e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = t.item()
        # UNet step latency: one graph replay (CFG-doubled UNet + guidance + DDIM update), median of 50
        st = next(iter(pipe._loops.values()))
        lat_ms = []
        for _ in range(60):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st.step.zero_()
            e0.record(); st.graph.replay(); e1.record()
            torch.cuda.synchronize()
            lat_ms.append(e0.elapsed_time(e1))
        unet_step_ms = statistics.median(lat_ms[10:])
        # tail share
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lat = pipe.denoise(lat_d, pos_d, neg_d, 1, GUIDANCE)
        torch.cuda.synchronize()
        e0.record()
        pipe.latents_to_waveform(lat)
        e1.record(); torch.cuda.synchronize()
        tail_ms = e0.elapsed_time(e1)
        by, launches_per_step = (profile_kernels(pipe, lat_d, pos_d, neg_d) if rank == 0 else ({}, 0))

    clips = BATCH * world * args.steps
    value = clips * CLIP_S / (ms_res / 1e3)
    e2e_value = clips * CLIP_S / e2e_s
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sustained, burst, hbm, how = peaks()
    traffic = None
    tp = ROOT / "profiles" / "conv_gemm_ncu_traffic.json"        # dram bytes of one profiled launch (ncu --set full)
    if tp.exists():
        traffic = json.loads(tp.read_text())
    cg = by.get("b200_conv_gemm", {"ms": 0.0, "calls": 0, "flops": 0.0})
    total_ms = sum(d["ms"] for d in by.values()) or 1.0
    achieved = cg["flops"] / (cg["ms"] / 1e3) / 1e12 if cg["ms"] else 0.0
    unet_flops = 101.74e9 * 2 * BATCH
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": CONFIG,
        "unet_step_ms": unet_step_ms,
        "unet_step_tflops": unet_flops / (unet_step_ms / 1e3) / 1e12,
        "tail_ms_per_batch": tail_ms,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(pos_h.numel() * 4 + neg_h.numel() * 4 + lat_h.numel() * 4),
                "d2h_bytes_per_step": int(audio.nbytes)},
        "gpu_launches": int(launches_per_step * STEPS_DDIM * args.steps),
        "launches_per_denoise_step": int(launches_per_step),
        "clocks": sampler.result(),
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM: all conv/linear layers)",
                     "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                     "peak_source": f"{how} bf16_tflops_sustained (kernel timed inside a long step)",
                     "how": f"sum of algorithmic FLOPs of the step's {cg['calls']} b200_conv_gemm launches (every conv / linear layer "
                            "except the LoRA-adapted projections, which run in b200_linear_lora) / sum of their per-launch "
                            "device times (each launch replayed 8x in its own CUDA graph, CUDA events on the launching stream)",
                     "launches_per_step": cg["calls"], "ms_per_step": cg["ms"], "share_of_step": cg["ms"] / total_ms,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE profiled launch of this kernel (ncu --set full);
                     # which launch, its algorithmic bytes and the other counters are in traffic_detail
                     "traffic": (traffic["dram_bytes_read"] + traffic["dram_bytes_write"]) if traffic else None,
                     "traffic_detail": traffic,
                     "whole_step": {"achieved": unet_flops / (unet_step_ms / 1e3) / 1e12, "frac": unet_flops / (unet_step_ms / 1e3) / 1e12 / sustained}},
        "kernel_breakdown_ms": {k: {"ms": round(v["ms"], 4), "calls": v["calls"],
                                    "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["flops"] and v["ms"] else None}
                                for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])},
    }
    if world == 1 and STEPS_DDIM >= 8:
        cb = cpu_reference_arm(4, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        try:
            line["gpu_eager_baseline"] = gpu_eager_arm(device)
        except Exception as exc:                         # noqa: BLE001 -- a baseline, never fatal for the bench line
            line["gpu_eager_baseline"] = {"unavailable": repr(exc)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
