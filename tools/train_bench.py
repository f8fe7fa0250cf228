#!/usr/bin/env python
"""LoRA fine-tuning step benchmark (BASELINE.json config 4): AudioLDM-S, rank-8 LoRA on attn q/k/v/out, frozen base,
synthetic latents [32, 8, 256, 16] per GPU, fp32 master LoRA weights, bf16 kernels, AdamW, NCCL all-reduce of the flat
LoRA-gradient arena when launched under torchrun.

    python tools/train_bench.py [--batch 32] [--steps 10] [--warmup 3] [--eager]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...

Prints one JSON line (rank 0): samples/s over all ranks, ms/step (max over ranks, CUDA events), model TFLOP/s.
Algorithmic FLOPs per sample: forward 103.92 G (SURVEY.md App. F, S train 256x16 r8); backward = dgrad of everything
downstream of the first adapted attention layer (forward minus the 6.7 % prefix minus SDPA) + 2.5 x SDPA
(5 instead of 2 matrix products) = 114.7 G; total 218.6 GFLOP/sample.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import audioldm_with_lora_b200 as b2  # noqa: E402
from audioldm_with_lora_b200 import synthetic  # noqa: E402
from audioldm_with_lora_b200.train import LoraTrainer  # noqa: E402

FLOP_PER_SAMPLE = 218.6e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--rank", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--eager", action="store_true", help="launch kernel by kernel from Python instead of one CUDA graph")
    ap.add_argument("--profile", action="store_true", help="per-C-ABI-call device time of one eager step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=dev)
    unet.load_state_dict(synthetic.random_lora_state_dict(cfg, args.rank, fmt="peft"), strict=False)
    trainer = LoraTrainer(unet, num_training_steps=97000 * world)
    g = torch.Generator().manual_seed(100 + rank)
    nb, h = args.batch, args.height
    lat = torch.randn(nb, 8, h, 16, generator=g).pin_memory()
    noise = torch.randn(nb, 8, h, 16, generator=g).pin_memory()
    t = torch.randint(0, 1000, (nb,), generator=g).pin_memory()
    emb = synthetic.clap_embeddings(nb * world)[0][rank * nb:(rank + 1) * nb].contiguous().pin_memory()
    step = trainer.train_step if args.eager else trainer.train_step_graphed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(args.warmup):
        losses.append(step(lat, noise, t, emb))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses.append(step(lat, noise, t, emb).clone())
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
    by = {}
    if args.profile and rank == 0:
        from audioldm_with_lora_b200 import _lib
        _lib.PROFILE = []
        trainer.train_step(lat, noise, t, emb)
        torch.cuda.synchronize()
        rec, _lib.PROFILE = _lib.PROFILE, None
        for name, a, b, info, _ in rec:
            d = by.setdefault(name, {"ms": 0.0, "calls": 0})
            d["ms"] += a.elapsed_time(b); d["calls"] += 1
        by = {k: {"ms": round(v["ms"], 3), "calls": v["calls"]} for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])}
    if rank == 0:
        ls = [float(x) for x in losses]
        line = {"metric": "lora_finetune_samples_per_sec", "value": nb * world / (ms / 1e3), "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "dtype": "bf16 (fp32 master LoRA weights / grads / AdamW)", "data": "synthetic",
                "config": {"workload": f"AudioLDM-S LoRA fine-tuning step, rank-{args.rank} q/k/v/out, batch {nb}/GPU, latents "
                                       f"{h}x16, frozen base, AdamW, {'eager launches' if args.eager else 'one CUDA graph per step'}",
                           "allreduce": f"NCCL sum over {world} ranks of the flat fp32 LoRA-grad arena ({trainer.numel * 4 / 1e6:.2f} MB)"
                           if world > 1 else "none (1 rank)"},
                "model_tflops": FLOP_PER_SAMPLE * nb / (ms / 1e3) / 1e12, "loss_first": ls[0], "loss_last": ls[-1],
                "arena_peak_gb": getattr(trainer, "arena_peak", 0) / 1e9, "eager_kernel_breakdown_ms": by}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
