"""CPU tests of the fine-tuning step's HOST logic (kernels emulated by tests/fake_ops.py): flat-arena layout, the
LoRA refresh table, the backward walk over the UNet graph, DDP gradient averaging over a 2-rank gloo group -- against
the fp32 autograd oracle (oracle/train_ref.py)."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
TINY = (64, 128, 192, 256)


def _install_fakes():
    from audioldm_with_lora_b200 import ops
    from tests import fake_ops
    for name in ("conv_gemm", "groupnorm_silu", "layernorm", "attention", "time_class_embed", "pack_nchw_to_nhwc",
                 "unpack_nhwc_to_nchw", "upsample_nearest", "sampler_step", "set_sm_budget", "gn_stat_slabs", "groupnorm_apply", "linear_lora_ok", "linear_lora", "linear_ln", "linear_stats", "add_noise", "adamw_flat",
                 "mse_partial") + fake_ops.TRAIN_OPS:
        setattr(ops, name, getattr(fake_ops, name))


def _tiny(rank=4, targets=("to_q", "to_k", "to_v", "to_out.0"), **kw):
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.arch import UNetConfig
    from audioldm_with_lora_b200.train import LoraTrainer
    from oracle import train_ref, unet_ref
    cfg = UNetConfig("tiny", TINY)
    sd = synthetic.random_unet_state_dict(cfg, seed=0)
    lsd = synthetic.random_lora_state_dict(cfg, rank, targets=targets, fmt="peft")
    unet = b2.UNet2DConditionModel(cfg, sd, device="cpu")
    unet.load_state_dict(lsd, strict=False)
    ad = b2.parse_lora_state_dict(lsd)
    spec = unet_ref.UNetSpec(block_out_channels=TINY, time_proj_dim=TINY[0])
    trainer = LoraTrainer(unet, **kw)
    ref = train_ref.TrainRef(sd, spec, {k: (e.A, e.B, e.alpha) for k, e in ad.items()}, **kw)
    return unet, trainer, ref


def _batch(nb, h, seed):
    from audioldm_with_lora_b200 import synthetic
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(nb, 8, h, 16, generator=g), torch.randn(nb, 8, h, 16, generator=g),
            torch.randint(0, 1000, (nb,), generator=g), synthetic.clap_embeddings(nb)[0])


def _flat(ref, trainer, what):
    out = torch.zeros(trainer.numel)
    for p, s in trainer.slots.items():
        A, B, _ = ref.params[p]
        a, b = (A.grad, B.grad) if what == "grad" else (A.detach(), B.detach())
        out[s.off_a: s.off_a + s.r * s.c] = a.reshape(-1)
        out[s.off_b: s.off_b + s.r * s.c] = b.reshape(-1)
    return out


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def test_flat_layout_and_refresh_table(fake_kernels):
    """The refresh kernel's descriptor table must reproduce exactly what the engine packs from the same adapters
    (forward operands), and the transposed counterparts (backward operands)."""
    unet, trainer, _ = _tiny(rank=4, targets=("to_q", "to_v", "to_out.0"))
    n_ad = len(trainer.slots)
    assert trainer.numel == sum(2 * s.r * s.c for s in trainer.slots.values())
    offs = sorted((s.off_a, s.off_b) for s in trainer.slots.values())
    assert offs[0][0] == 0 and all(b == a + s for (a, b), s in zip(offs, [trainer.slots[p].r * trainer.slots[p].c
                                                                    for p in trainer.slots]))
    nb, h, w = 2, 16, 16
    plan = unet.engine._plan(nb, h, w)
    before = {k: v.w.clone() for k, v in plan["W"].items()}
    tp = trainer._tplan(nb, h, w)
    assert tp["refresh_n"] == 4 * n_ad
    trainer.refresh(nb, h, w)                           # flat_p was initialised from the same adapters: a no-op
    for k, v in plan["W"].items():
        assert torch.equal(v.w, before[k]), k
    # backward operands: K segment = A^T, down-projection = (s B)^T
    p = next(iter(trainer.slots)).rsplit(".to_", 1)[0]
    c = trainer.slots[p + ".to_q"].c
    A_q, B_q = trainer.param_views()[p + ".to_q"]
    wq = tp["Wb"][p + ".bwd_qkv"].w
    assert torch.equal(wq[:c, 3 * c: 3 * c + 4], A_q.t().to(torch.bfloat16))
    assert torch.equal(tp["Wb"][p + ".bwd_down_qkv"].w[:4, :c], B_q.t().to(torch.bfloat16))
    # a parameter update flows into every packed operand
    trainer.flat_p.mul_(2.0)
    old = wq[:c, 3 * c: 3 * c + 4].clone()
    assert not torch.equal(old, A_q.t().to(torch.bfloat16))          # A_q is a view of flat_p: already doubled
    trainer.refresh(nb, h, w)
    assert torch.equal(wq[:c, 3 * c: 3 * c + 4], A_q.t().to(torch.bfloat16))
    assert torch.equal(plan["W"][p + ".lora_down_qkv"].w[:4], A_q.to(torch.bfloat16))


@pytest.mark.parametrize("rank,targets,nb,h", [(4, ("to_q", "to_k", "to_v", "to_out.0"), 2, 16),
                                               (2, ("to_q", "to_v"), 1, 12)])
def test_backward_walk_matches_autograd_oracle(fake_kernels, rank, targets, nb, h):
    unet, trainer, ref = _tiny(rank, targets)
    lat, noise, t, emb = _batch(nb, h, seed=5)
    loss_ref = ref.loss_and_grads(lat, noise, t, emb)
    noisy = ref.noise_sched.add_noise(lat, noise, t)
    loss = trainer.forward_backward(noisy, t, emb, noise)
    assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < 1e-2
    g, g_ref = trainer.flat_g, _flat(ref, trainer, "grad")
    assert rel(g, g_ref) < 6e-2, rel(g, g_ref)        # activations / gradients are rounded to bf16 between "kernels"
    assert not trainer.arena.live                      # arena fully reclaimed


def test_train_step_adamw_and_polynomial_lr(fake_kernels):
    kw = dict(lr=1e-3, weight_decay=1e-2, num_training_steps=4)
    unet, trainer, ref = _tiny(4, **kw)
    p0 = trainer.flat_p.clone()
    for step in range(3):
        lat, noise, t, emb = _batch(2, 16, seed=30 + step)
        trainer.train_step(lat, noise, t, emb)
        ref.train_step(lat, noise, t, emb)
        assert abs(trainer.current_lr() - ref.opt.param_groups[0]["lr"]) < 1e-12
    upd, upd_ref = trainer.flat_p - p0, _flat(ref, trainer, "param") - p0
    assert torch.dot(upd, upd_ref) / (upd.norm() * upd_ref.norm()) > 0.97
    from audioldm_with_lora_b200.train import polynomial_lr
    assert polynomial_lr(0, 1e-5, 100) == pytest.approx(1e-5)
    assert polynomial_lr(100, 1e-5, 100) == pytest.approx(1e-7)
    assert polynomial_lr(1000, 1e-5, 100) == pytest.approx(1e-7)


def test_eval_checkpoint_and_pipeline_follow_an_external_optimizer(fake_kernels, tmp_path):
    """the reference-shaped loop (torch AdamW over the peft Parameters, loss.backward()) changes the flat
    arena behind the model's back.  Every later eval forward, `get_peft_model_state_dict`, `save_attn_procs` and the
    validation pipeline must see the NEWEST parameters -- train, eval, train, eval."""
    import torch.nn.functional as F
    import audioldm_with_lora_b200 as b2
    from safetensors.torch import load_file
    unet, _, _ = _tiny(rank=4)
    unet.requires_grad_(False)
    tr = unet.lora_trainer()
    params = [p for p in unet.parameters() if p.requires_grad]
    assert len(params) == 2 * len(tr.slots)
    opt = torch.optim.AdamW(params, lr=1e-2)
    pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), use_cuda_graph=False)
    lat_e, _, _, emb_e = _batch(1, 16, seed=5)
    key = next(k for k in b2.get_peft_model_state_dict(unet) if "lora_B" in k)
    slot = tr.slots[key.split("base_model.model.")[1].split(".lora_B")[0]]
    seen_eps, seen_sd, seen_pipe = [], [], []
    for rnd in range(2):
        unet.train()
        lat, noise, t, emb = _batch(2, 16, seed=60 + rnd)
        pred = unet(lat, t, encoder_hidden_states=None, class_labels=emb, cross_attention_kwargs={"scale": 1.0},
                    return_dict=False)[0]
        F.mse_loss(pred.float(), noise.float()).backward()
        opt.step(); opt.zero_grad()
        unet.eval()
        with torch.no_grad():
            seen_eps.append(unet(lat_e, 300, class_labels=emb_e, return_dict=False)[0].clone())
        seen_sd.append(b2.get_peft_model_state_dict(unet)[key].clone())
        f = unet.save_attn_procs(tmp_path / f"ckpt{rnd}")
        seen_pipe.append(pipe.denoise(lat_e.clone(), emb_e, emb_e.roll(1, 1), 2, 2.5).clone())
        on_disk = load_file(str(f))
        assert torch.equal(seen_sd[-1].reshape(-1), tr.flat_p.detach()[slot.off_b: slot.off_b + slot.r * slot.c])
        assert any(v.numel() == seen_sd[-1].numel() and torch.equal(v.reshape(-1), seen_sd[-1].reshape(-1))
                   for v in on_disk.values())
    assert not torch.equal(seen_sd[0], seen_sd[1])
    assert not torch.equal(seen_eps[0], seen_eps[1])
    assert not torch.equal(seen_pipe[0], seen_pipe[1])


def test_repack_in_place_keeps_device_pointers(fake_kernels):
    """`engine.set_lora` / `set_lora_scale` on unchanged shapes overwrites the packed tensors in place, so
    pointers held by captured CUDA graphs and by the trainer's refresh table stay valid and `weights_version` (what
    `train_step_graphed` and the pipeline key their graphs on) does not move; a wider LoRA segment moves both."""
    from audioldm_with_lora_b200 import synthetic
    unet, trainer, _ = _tiny(rank=4)
    eng = unet.engine
    plan = eng._plan(2, 16, 16)
    ptrs = {k: v.w.data_ptr() for k, v in plan["W"].items()}
    v0 = eng.weights_version
    p = next(k for k in plan["W"] if k.endswith("attn1.qkv"))
    before = plan["W"][p].w.clone()
    eng.set_lora_scale(0.5)
    assert eng.weights_version == v0 and {k: v.w.data_ptr() for k, v in plan["W"].items()} == ptrs
    assert not torch.equal(plan["W"][p].w, before)              # ...but the values did change
    eng.set_lora_scale(1.0)
    assert torch.equal(plan["W"][p].w, before)
    lsd = synthetic.random_lora_state_dict(unet.cfg, 40, fmt="peft")       # q,k,v: r_total = 120 > 64: wider K segment
    unet.load_state_dict(lsd, strict=False)
    assert eng.weights_version > v0


def test_abandoned_grad_mode_forward_fails_loudly(fake_kernels):
    """two grad-mode forwards before one backward share one activation arena -- the older one's backward
    must raise instead of reading overwritten activations."""
    unet, _, _ = _tiny(rank=4)
    unet.requires_grad_(False)
    unet.lora_trainer()
    unet.train()
    lat, noise, t, emb = _batch(1, 16, seed=3)
    first = unet(lat, t, class_labels=emb, return_dict=False)[0]
    second = unet(lat, t, class_labels=emb, return_dict=False)[0]
    second.float().pow(2).mean().backward()
    with pytest.raises(RuntimeError, match="activations of this forward are gone"):
        first.float().pow(2).mean().backward()


# ----------------------------------------------------------------------------------------- N > 1: DDP over gloo
def _ddp_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    _install_fakes()
    kw = dict(lr=1e-3, weight_decay=0.0)
    unet, trainer, _ = _tiny(4, **kw)
    lat, noise, t, emb = _batch(1, 16, seed=70 + rank)          # each rank its own micro-batch
    trainer.train_step(lat, noise, t, emb)                       # forward/backward, all-reduce, AdamW with 1/world
    torch.save({"p": trainer.flat_p.clone(), "g": trainer.flat_g.clone()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ddp_step_matches_oracle_gradient_average(tmp_path, fake_kernels):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_ddp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["g"], r1["g"]) and torch.equal(r0["p"], r1["p"])       # replicas stay in lock-step
    kw = dict(lr=1e-3, weight_decay=0.0)
    _, trainer, ref0 = _tiny(4, **kw)
    _, _, ref1 = _tiny(4, **kw)
    for rk, ref in enumerate((ref0, ref1)):
        ref.loss_and_grads(*_batch(1, 16, seed=70 + rk))
    ref0.average_grads_with([ref1])
    g_mean = _flat(ref0, trainer, "grad")
    assert rel(r0["g"] / world, g_mean) < 6e-2                   # the arena holds the SUM; AdamW applies 1/world
    p0 = trainer.flat_p.clone()
    ref0.optimizer_step()
    upd, upd_ref = r0["p"] - p0, _flat(ref0, trainer, "param") - p0
    assert torch.dot(upd, upd_ref) / (upd.norm() * upd_ref.norm()) > 0.97
