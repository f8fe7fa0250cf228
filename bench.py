#!/usr/bin/env python
"""bench.py -- generated audio-sec/sec of the LoRA-adapted AudioLDM sampling path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): AudioLDM-S architecture (random-init, seed 0) + rank-8 LoRA on
attn q/k/v/out, batch 8 prompts per GPU, 200 DDIM steps, CFG 2.5, 10 s clips, bf16 kernels with fp32
accumulation, synthetic L2-normalised CLAP embeddings.  One bench "step" = one batch of 8 clips
through the whole path (200 x {CFG-doubled UNet, guidance, DDIM update} + VAE decode + HiFi-GAN vocoder, all of it on the
hand-written sm_100a kernels).

  value  : B*10 s*K / device time, inputs already resident in HBM (denoise loop + the VAE / vocoder tail, both replayed from CUDA graphs)
  e2e    : the same through the public call `AudioLDMPipeline.__call__(prompt_embeds=<host tensors>, ...)`
           returning host numpy audio (H2D of embeddings/latents and D2H of waveforms inside the timing)
  roofline: the implicit-GEMM kernel (all conv / linear layers, ~89 % of the step's FLOPs) timed per
           launch with CUDA events on the launching stream; algorithmic FLOPs / time vs measured bf16 peak
  cpu_baseline: the torch-eager fp32 oracle (restating the reference's diffusers path) on the host cores,
           bounded sample, extrapolated to the same 200-step / 10 s clip.
  roofline_hbm: the HBM-bound kernels of the step (GroupNorm+SiLU, LayerNorm, sampler update, up-sampling): algorithmic
           bytes / per-launch device time vs the measured copy bandwidth.
--impl reference prints that CPU arm as its own line (rank 0 only).

After the c2 headline the same run measures BASELINE.json's other configurations as short extra legs and reports them
under `extra_configs` (they are NOT the `value`; --legs c2 skips them):
  c3       : AudioLDM-S + rank-16 LoRA, 64 prompts SHARDED over the N GPUs (strong scaling), 200 DDIM steps, 10 s clips,
             through the public pipeline call; audio-s/s over all ranks
  c4_train : LoRA fine-tuning step, batch 32/GPU, latents 256x16, one CUDA graph per step + NCCL all-reduce of the flat
             LoRA-gradient arena + fused AdamW; ms/step, samples/s, the all-reduce alone in microseconds
  c5       : AudioLDM-L + rank-32 LoRA, 30 s clips (750x16 latents), batch 16/GPU (UNet batch 32): ms per denoising step
             against the 29.2 ms tensor-bound floor
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
            "tail_s": tail, "ms_per_bench_step": BATCH * clip_time * 1e3}


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


def _max_over_ranks(x: float, device, world: int) -> float:
    if world <= 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def _barrier(world: int) -> None:
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def leg_c3(device, rank: int, world: int) -> dict:
    """BASELINE configs[2]: S + rank-16 LoRA, 64 prompts sharded over the ranks by global prompt index (strong scaling:
    the job is fixed, per-GPU batch = 64 / N), 200 DDIM steps, CFG 2.5, 10 s clips, through the public pipeline call
    with host tensors (H2D + D2H inside the timing).  One warm-up call (graph capture), one timed call."""
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic, tail
    from audioldm_with_lora_b200.sharding import shard_prompts
    total = 64
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=device)
    unet.load_state_dict(synthetic.random_lora_state_dict(cfg, 16, fmt="peft"), strict=False)
    pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), vae=tail.random_vae_decoder(7), vocoder=tail.build_vocoder(0))
    pos, neg = synthetic.clap_embeddings(total)
    pos, neg, owned = shard_prompts(pos, neg, rank, world)
    pos, neg = pos.contiguous().pin_memory(), neg.contiguous().pin_memory()
    lat = synthetic.initial_latents(len(owned), int(CLIP_S / 0.01) // 4, first_index=owned[0]).pin_memory()

    def call(steps):
        return pipe(prompt_embeds=pos, negative_prompt_embeds=neg, latents=lat, audio_length_in_s=CLIP_S,
                    num_inference_steps=steps, guidance_scale=GUIDANCE).audios

    with torch.no_grad():
        call(3)                                   # packs weights, captures the step graph and the tail graph
        _barrier(world)
        t0 = time.perf_counter()
        audio = call(STEPS_DDIM)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
    sec = _max_over_ranks(sec, device, world)
    ok = bool(audio.shape == (len(owned), int(CLIP_S * 16000)))
    return {"workload": "AudioLDM-S + rank-16 LoRA q/k/v/out, 64 prompts sharded over the GPUs, 200 DDIM steps, CFG 2.5, 10 s clips",
            "metric": METRIC, "value": total * CLIP_S / sec, "unit": UNIT, "scaling": "strong", "n_gpus": world,
            "prompts_per_gpu": len(owned), "seconds_per_job": sec, "measured": "1 public pipeline call per rank with host "
            "tensors (max over ranks), after a 3-step warm-up call", "output_ok": ok, "collectives_on_data_path": 0}


def leg_c4_train(device, rank: int, world: int, steps: int = 10, warmup: int = 3) -> dict:
    """BASELINE configs[3]: LoRA fine-tuning step (train_audioldm_lora.py:499-565), frozen base, rank-8 adapters, synthetic
    latents [32, 8, 256, 16] per GPU copied from pinned host memory every step, {refresh, add_noise, forward, MSE,
    backward} as one CUDA graph, then ONE NCCL all-reduce of the flat fp32 LoRA-gradient arena and the fused AdamW."""
    import torch.distributed as dist
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.train import LoraTrainer
    nb, h = 32, 256
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=device)
    unet.load_state_dict(synthetic.random_lora_state_dict(cfg, RANK_LORA, fmt="peft"), strict=False)
    trainer = LoraTrainer(unet, num_training_steps=97000)
    g = torch.Generator().manual_seed(100 + rank)
    lat = torch.randn(nb, 8, h, 16, generator=g).pin_memory()
    noise = torch.randn(nb, 8, h, 16, generator=g).pin_memory()
    t = torch.randint(0, 1000, (nb,), generator=g).pin_memory()
    emb = synthetic.clap_embeddings(nb * world)[0][rank * nb:(rank + 1) * nb].contiguous().pin_memory()
    losses = []
    for _ in range(warmup):
        losses.append(float(trainer.train_step_graphed(lat, noise, t, emb)))
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = trainer.train_step_graphed(lat, noise, t, emb)
    e1.record()
    _barrier(world)
    losses.append(float(loss))
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, device, world)
    ar_us = None
    if world > 1:
        scratch = torch.zeros_like(trainer.flat_g)
        for _ in range(3):
            dist.all_reduce(scratch)
        _barrier(world)
        e0.record()
        for _ in range(20):
            dist.all_reduce(scratch)
        e1.record()
        torch.cuda.synchronize()
        ar_us = _max_over_ranks(e0.elapsed_time(e1) / 20 * 1e3, device, world)
    flop_per_sample = 218.6e9            # fwd 103.92 G + dgrad downstream of the first adapted layer (tools/train_bench.py)
    return {"workload": "AudioLDM-S LoRA fine-tuning step, rank-8 q/k/v/out, frozen base, batch 32/GPU, latents 256x16, bf16 "
                        "kernels, fp32 master LoRA weights / grads / AdamW", "metric": "lora_finetune_samples_per_sec",
            "value": nb * world / (ms / 1e3), "unit": "samples/s", "scaling": "weak", "n_gpus": world, "ms_per_step": ms,
            "steps": steps, "warmup": warmup, "model_tflops_per_gpu": flop_per_sample * nb / (ms / 1e3) / 1e12,
            "allreduce": {"collective": "NCCL all_reduce(SUM) of the flat fp32 LoRA-grad arena, 1 per step", "ranks": world,
                          "bytes": int(trainer.numel * 4), "us_alone": ar_us},
            "h2d_bytes_per_step": int(lat.numel() * 8 + emb.numel() * 4 + t.numel() * 8),
            "loss_first": losses[0], "loss_last": losses[-1]}


def leg_c5(device, rank: int, world: int, reps: int = 6) -> dict:
    """BASELINE configs[4]: AudioLDM-L (739 M) + rank-32 LoRA, 30 s clips (750x16 latents), batch 16 per GPU (UNet batch 32
    with CFG).  One denoising step = 41.24 TFLOP: a 200-step run is ~25 s per batch, so this leg times the step itself
    (graph replays of {UNet, guidance, DDIM update}) and reports ms/step against the 29.2 ms tensor-bound floor."""
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic
    nb, h = 16, 750
    cfg = b2.CONFIGS["L"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=device)
    unet.load_attn_procs(synthetic.random_lora_state_dict(cfg, 32, fmt="diffusers"))
    pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler())
    pos, neg = synthetic.clap_embeddings(nb * world)
    pos, neg = pos[rank * nb:(rank + 1) * nb].to(device), neg[rank * nb:(rank + 1) * nb].to(device)
    lat = synthetic.initial_latents(nb, h, first_index=rank * nb).to(device)
    with torch.no_grad():
        out = pipe.denoise(lat, pos, neg, 2, GUIDANCE)          # capture + warm
        st = next(iter(pipe._loops.values()))
        times = []
        for _ in range(reps):
            st.step.zero_()
            _barrier(world)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); st.graph.replay(); e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
    ms = _max_over_ranks(statistics.median(times[1:]), device, world)
    sustained = peaks()[0]
    flop = 1288.77e9 * 2 * nb
    return {"workload": "AudioLDM-L-full arch (random-init) + rank-32 LoRA q/k/v/out, 30 s clips (750x16 latents), batch 16/GPU, CFG 2.5",
            "metric": "unet_step_ms", "value": ms, "unit": "ms per denoising step (UNet batch 32 + guidance + DDIM update)",
            "higher_is_better": False, "scaling": "weak", "n_gpus": world, "tflops_per_gpu": flop / (ms / 1e3) / 1e12,
            "frac_of_sustained_bf16_peak": flop / (ms / 1e3) / 1e12 / sustained, "tensor_bound_floor_ms": flop / sustained / 1e9,
            "implied_audio_s_per_s_loop_only": nb * world * 30.0 / (200 * ms / 1e3), "finite": bool(torch.isfinite(out).all()),
            "measured": f"median of {reps - 1} graph replays, max over ranks"}


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
        d = by.setdefault(name, {"ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
        d["ms"] += e0.elapsed_time(e1) / reps; d["calls"] += 1
        d["flops"] += (info or {}).get("flops", 0.0)
        d["bytes"] += (info or {}).get("bytes", 0.0)
    torch.cuda.current_stream().wait_stream(side)
    return by, launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--legs", default=os.environ.get("B200_BENCH_LEGS", "c2,c3,c4,c5"),
                    help="comma list; c2 (the headline) always runs, c3 / c4 / c5 are the extra BASELINE configurations")
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
        cfg_ref = dict(CONFIG)
        cfg_ref["reference_sample"] = ("TIMED: 1 prompt (UNet batch 2), %d CFG denoising steps + the VAE/vocoder tail once, fp32, on %d "
                                       "host cores; EXTRAPOLATED to the workload above (8 clips x (200 steps + tail), clips in "
                                       "sequence: the CPU gains nothing from batching)" % (max(args.steps, 1), r["cores"]))
        # ms_per_step has ONE meaning in both arms: time of one bench step = one batch of 8 clips through the whole path
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_bench_step"], "ms_per_denoise_step_unet_batch2": r["ms_per_denoise_step"],
                "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_ref, "host_cores": r["cores"],
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

    pipe_vae, pipe_voc, tail_launches = type(pipe.vae).__name__, type(pipe.vocoder).__name__, pipe.tail_launches
    clips = BATCH * world * args.steps
    value = clips * CLIP_S / (ms_res / 1e3)
    e2e_value = clips * CLIP_S / e2e_s
    # ---- BASELINE.json's other configurations as short extra legs (all ranks take part; rank 0 reports)
    legs = {x.strip() for x in args.legs.split(",")}
    extra = {}
    if STEPS_DDIM == 200:
        del pipe, st
        torch.cuda.empty_cache()
        for name, fn in (("c3", leg_c3), ("c4", leg_c4_train), ("c5", leg_c5)):
            if name not in legs:
                continue
            try:
                extra[fn.__name__[4:]] = fn(device, rank, world)
            except Exception as exc:                     # noqa: BLE001 -- an extra leg never takes the headline line down
                extra[fn.__name__[4:]] = {"failed": repr(exc)[:300]}
            torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sustained, burst, hbm, how = peaks()
    traffic = None
    tp = ROOT / "profiles" / "conv_gemm_ncu_traffic.json"        # dram bytes of one profiled launch (ncu --set full)
    if tp.exists():
        traffic = json.loads(tp.read_text())
    cg = {"ms": 0.0, "calls": 0, "flops": 0.0}
    for k_ in ("b200_conv_gemm", "b200_conv_gemm_gnstat"):      # the same kernel; the second entry point also leaves GroupNorm partials
        for f_ in cg:
            cg[f_] += by.get(k_, {}).get(f_, 0)
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
        "tail": {"vae_decoder": pipe_vae, "vocoder": pipe_voc, "launches_per_batch": int(tail_launches),
                 "note": "B200VaeDecoder / B200HifiGan = the same sm_100a kernels as the UNet (SURVEY 8f items 1-2); torch modules = reference path"},
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(pos_h.numel() * 4 + neg_h.numel() * 4 + lat_h.numel() * 4),
                "d2h_bytes_per_step": int(audio.nbytes)},
        "gpu_launches": int((launches_per_step * STEPS_DDIM + tail_launches) * args.steps),
        "launches_per_denoise_step": int(launches_per_step),
        "clocks": sampler.result(),
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM: all conv/linear layers)",
                     "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                     "peak_source": f"{how} bf16_tflops_sustained (kernel timed inside a long step)",
                     "how": f"sum of algorithmic FLOPs of the step's {cg['calls']} b200_conv_gemm / b200_conv_gemm_gnstat launches (every conv / linear layer "
                            "except the LoRA-adapted projections, which run in b200_linear_lora) / sum of their per-launch "
                            "device times (each launch replayed 8x in its own CUDA graph, CUDA events on the launching stream)",
                     "launches_per_step": cg["calls"], "ms_per_step": cg["ms"], "share_of_step": cg["ms"] / total_ms,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE profiled launch of this kernel (ncu --set full);
                     # which launch, its algorithmic bytes and the other counters are in traffic_detail
                     "traffic": (traffic["dram_bytes_read"] + traffic["dram_bytes_write"]) if traffic else None,
                     "traffic_detail": traffic,
                     "whole_step": {"achieved": unet_flops / (unet_step_ms / 1e3) / 1e12, "frac": unet_flops / (unet_step_ms / 1e3) / 1e12 / sustained}},
        "kernel_breakdown_ms": {k: {"ms": round(v["ms"], 4), "calls": v["calls"],
                                    "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["flops"] and v["ms"] else None,
                                    "gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) if v["bytes"] and v["ms"] else None}
                                for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])},
        "kernel_timing": "per launch, each replayed 8x back to back in its own CUDA graph: operands L2-WARM, isolated from its "
                         "neighbours in the step (the sum over launches reproduces unet_step_ms to ~2 %)",
        # the norm / elementwise kernels against the measured copy bandwidth (north_star: "achieved HBM GB/s ... against peak")
        "roofline_hbm": [{"kernel": k, "bound": "hbm", "launches_per_step": by[k]["calls"], "bytes_per_step": by[k]["bytes"],
                          "ms_per_step": round(by[k]["ms"], 4), "achieved": round(by[k]["bytes"] / (by[k]["ms"] / 1e3) / 1e9, 1),
                          "peak": hbm, "unit": "GB/s", "frac": round(by[k]["bytes"] / (by[k]["ms"] / 1e3) / 1e9 / hbm, 4),
                          "bytes_are": "algorithmic: one read + one write of the tensor (GroupNorm's second read is an L2 hit)"}
                         for k in ("b200_groupnorm_apply", "b200_groupnorm_silu", "b200_layernorm", "b200_sampler_step", "b200_upsample_nearest")
                         if k in by and by[k]["bytes"] and by[k]["ms"]],
        "extra_configs": extra,
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
