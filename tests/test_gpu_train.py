"""GPU parity tests (pytest -m gpu) of the LoRA fine-tuning step (BASELINE config 4) against
the fp32 CPU oracle (oracle/train_ref.py: torch autograd + torch.optim.AdamW, i.e. what the reference runs).

Tolerances (bf16 kernels with fp32 accumulation vs the fp32 oracle, stated here as BASELINE.json asks):
  loss              relative error <= 1e-2
  LoRA gradients    rel-L2 over the whole flat arena <= 8e-2; per adapter matrix |g - g_ref| <= 0.2 * max(|g_ref|, 5 % of
                    the largest adapter-gradient norm)  (matrices with near-zero gradients are cancellation noise)
  parameters        after 3 optimizer steps, update direction cosine >= 0.98 and rel-L2 of the update <= 1.5e-1
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOSS_TOL = 1e-2
GRAD_TOL = 8e-2
GRAD_TOL_ADAPTER = 2e-1


@pytest.fixture(scope="module", autouse=True)
def _need_cuda_and_lib():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the hot path has no CPU fallback")
    from audioldm_with_lora_b200 import _lib
    _lib.load()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _setup(rank=8, targets=("to_q", "to_k", "to_v", "to_out.0"), alpha=None, **trainer_kw):
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.train import LoraTrainer
    from oracle import train_ref, unet_ref
    cfg = b2.CONFIGS["S"]
    sd = synthetic.random_unet_state_dict(cfg, seed=0)
    lsd = synthetic.random_lora_state_dict(cfg, rank, targets=targets, fmt="peft")
    unet = b2.UNet2DConditionModel(cfg, sd, device=DEV)
    unet.load_lora_state_dict(lsd, alpha=alpha)
    ad = b2.parse_lora_state_dict(lsd, alpha)
    trainer = LoraTrainer(unet, **trainer_kw)
    ref = train_ref.TrainRef(sd, unet_ref.ARCH_S, {k: (e.A, e.B, e.alpha) for k, e in ad.items()}, **trainer_kw)
    return unet, trainer, ref


def _batch(nb, h, seed=11):
    from audioldm_with_lora_b200 import synthetic
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(nb, 8, h, 16, generator=g)
    noise = torch.randn(nb, 8, h, 16, generator=g)
    t = torch.randint(0, 1000, (nb,), generator=g)
    emb, _ = synthetic.clap_embeddings(nb)
    return lat, noise, t, emb


def _flat(ref, trainer, what):
    """Oracle gradients / parameters in the trainer's flat layout."""
    out = torch.zeros(trainer.numel)
    for p, s in trainer.slots.items():
        A, B, _ = ref.params[p]
        a, b = (A.grad, B.grad) if what == "grad" else (A.detach(), B.detach())
        out[s.off_a: s.off_a + s.r * s.c] = a.reshape(-1)
        out[s.off_b: s.off_b + s.r * s.c] = b.reshape(-1)
    return out


@pytest.mark.parametrize("rank,targets,alpha,nb,h", [
    (8, ("to_q", "to_k", "to_v", "to_out.0"), None, 2, 32),     # config-4 adapters at a small latent
    (2, ("to_q", "to_v"), 2.0, 1, 24),                          # the reference's own LoraConfig (train:378-383)
    (16, ("to_q", "to_k", "to_v", "to_out.0"), 32.0, 2, 20),    # alpha = 2r, odd level sizes (20 -> 10 -> 5 -> 3)
])
def test_loss_and_lora_grads_match_oracle(rank, targets, alpha, nb, h):
    unet, trainer, ref = _setup(rank, targets, alpha)
    lat, noise, t, emb = _batch(nb, h)
    loss_ref = ref.loss_and_grads(lat, noise, t, emb)
    noisy = ref.noise_sched.add_noise(lat, noise, t)
    trainer.flat_g.zero_()
    loss = trainer.forward_backward(noisy.to(DEV), t.to(DEV), emb.to(DEV), noise.to(DEV))
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < LOSS_TOL
    g_ref = _flat(ref, trainer, "grad")
    g = trainer.flat_g.cpu()
    assert torch.isfinite(g).all()
    slices = [slice(off, off + s.r * s.c) for s in trainer.slots.values() for off in (s.off_a, s.off_b)]
    biggest = max(g_ref[sl].norm().item() for sl in slices)
    worst = max((g[sl] - g_ref[sl]).norm().item() / max(g_ref[sl].norm().item(), 0.05 * biggest) for sl in slices)
    assert rel(g, g_ref) < GRAD_TOL, f"flat LoRA-grad rel-L2 {rel(g, g_ref):.3e}"
    assert worst < GRAD_TOL_ADAPTER, f"worst adapter-matrix error {worst:.3e}"


def test_lora_grad_error_is_bf16_storage_not_arithmetic():
    """Where the 4-5e-2 against fp32 autograd comes from.  The oracle re-run with every inter-kernel tensor (and the frozen
    weights) ROUNDED TO BF16 -- values straight-through, gradients through hooks, all arithmetic still fp32
    (oracle/unet_ref.py::STORAGE_ROUND) -- moves away from the plain fp32 oracle by as much as the kernels do (measured on
    B200: 4.31e-2 for the rounding oracle, 4.37e-2 for the kernels, cosine 0.9992; the two differ from each other by 4.8e-2,
    i.e. two realisations of the same rounding noise): the error budget is bf16 STORAGE of activations / gradients through
    ~60 layers (what the reference's own `--mixed_precision bf16` runs incur), not the kernels' arithmetic.  Numbers go to
    gpurun_out/train_parity_report.json."""
    import json
    from pathlib import Path
    from oracle import unet_ref
    unet, trainer, ref = _setup(8, ("to_q", "to_k", "to_v", "to_out.0"), None)
    lat, noise, t, emb = _batch(2, 32)
    ref.loss_and_grads(lat, noise, t, emb)
    g_fp32 = _flat(ref, trainer, "grad").clone()
    # the same oracle with bf16 storage: weights as the kernels hold them, activations / gradients rounded between "kernels"
    sd_keep = ref.sd
    ref.sd = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v) for k, v in sd_keep.items()}
    unet_ref.STORAGE_ROUND = torch.bfloat16
    try:
        ref.loss_and_grads(lat, noise, t, emb)
    finally:
        unet_ref.STORAGE_ROUND = None
        ref.sd = sd_keep
    g_emul = _flat(ref, trainer, "grad").clone()
    noisy = ref.noise_sched.add_noise(lat, noise, t)
    trainer.flat_g.zero_()
    trainer.forward_backward(noisy.to(DEV), t.to(DEV), emb.to(DEV), noise.to(DEV))
    torch.cuda.synchronize()
    g = trainer.flat_g.cpu()
    rep = {"kernels_vs_fp32_oracle": rel(g, g_fp32), "kernels_vs_bf16_storage_oracle": rel(g, g_emul),
           "bf16_storage_oracle_vs_fp32_oracle": rel(g_emul, g_fp32),
           "cosine_kernels_fp32": torch.nn.functional.cosine_similarity(g, g_fp32, dim=0).item()}
    out = Path(__file__).resolve().parents[1] / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "train_parity_report.json").write_text(json.dumps(rep, indent=1))
    print(rep)
    assert rep["kernels_vs_fp32_oracle"] < GRAD_TOL
    # storage rounding alone explains an error of the same size ...
    assert rep["bf16_storage_oracle_vs_fp32_oracle"] > 0.4 * rep["kernels_vs_fp32_oracle"]
    # ... and the kernels are no further from the fp32 truth than 2x what storage rounding alone costs
    assert rep["kernels_vs_fp32_oracle"] < 2.0 * rep["bf16_storage_oracle_vs_fp32_oracle"] + 1e-2


def test_train_steps_match_oracle_adamw():
    kw = dict(lr=1e-3, weight_decay=1e-2, num_training_steps=10)
    unet, trainer, ref = _setup(8, **kw)
    p0 = trainer.flat_p.cpu().clone()
    losses, losses_ref = [], []
    for step in range(3):
        lat, noise, t, emb = _batch(2, 32, seed=20 + step)
        losses.append(trainer.train_step(lat, noise, t, emb).item())
        losses_ref.append(ref.train_step(lat, noise, t, emb).item())
    for a, b in zip(losses, losses_ref):
        assert abs(a - b) / b < 2e-2
    upd = trainer.flat_p.cpu() - p0
    upd_ref = _flat(ref, trainer, "param") - p0
    cos = torch.dot(upd, upd_ref) / (upd.norm() * upd_ref.norm())
    assert cos > 0.98, f"update cosine {cos:.4f}"
    assert rel(upd, upd_ref) < 1.5e-1
    # the trained adapters flow back into the inference engine: sampling-path forward == oracle with the new LoRA
    trainer.sync_to_model()
    from oracle import unet_ref
    lat, _, t, emb = _batch(2, 32, seed=99)
    eps = unet(lat.to(DEV), 500, class_labels=emb.to(DEV), return_dict=False)[0]
    with torch.no_grad():
        lora = unet_ref.LoraSet({p: (A.detach(), B.detach(), al) for p, (A, B, al) in ref.params.items()})
        eps_ref = unet_ref.unet_forward(ref.sd, unet_ref.ARCH_S, lat, 500, emb, lora=lora)
    assert rel(eps, eps_ref) < 3e-2


def test_zero_B_gives_zero_A_grad_and_full_size_step_is_finite():
    """Known-answer property (peft init: B = 0 => dA = 0, dB != 0) at the config-4 latent size (256 x 16), batch 2."""
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.lora import LoraConfig
    from audioldm_with_lora_b200.train import LoraTrainer
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=DEV)
    b2.get_peft_model(unet, LoraConfig(r=8, lora_alpha=8, target_modules=["to_q", "to_k", "to_v", "to_out.0"]))
    trainer = LoraTrainer(unet)
    lat, noise, t, emb = _batch(2, 256)
    noisy = lat * 0.7 + noise * 0.7
    trainer.flat_g.zero_()
    loss = trainer.forward_backward(noisy.to(DEV), t.to(DEV), emb.to(DEV), noise.to(DEV))
    g = trainer.grad_views()
    assert torch.isfinite(loss) and torch.isfinite(trainer.flat_g).all()
    a_norm = sum(float(a.norm()) for a, _ in g.values())
    b_norm = sum(float(b.norm()) for _, b in g.values())
    assert a_norm == 0.0 and b_norm > 0.0


def test_reference_shaped_training_loop_through_autograd():
    """The reference's own loop shape (train_audioldm_lora.py:373-403, 479, 539-565): get_peft_model, AdamW over
    filter(requires_grad), unet.train(), unet(...)[0], F.mse_loss, loss.backward(), optimizer.step() -- with torch's
    optimizer and autograd driving the B200 forward / backward through the autograd.Function seam."""
    import torch.nn.functional as F
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic
    from oracle import train_ref, unet_ref
    cfg = b2.CONFIGS["S"]
    sd = synthetic.random_unet_state_dict(cfg, seed=0)
    lsd = synthetic.random_lora_state_dict(cfg, 8, fmt="peft")
    unet = b2.UNet2DConditionModel(cfg, sd, device=DEV)
    unet.load_state_dict(lsd, strict=False)
    ad = b2.parse_lora_state_dict(lsd)
    kw = dict(lr=1e-3, weight_decay=1e-2)
    ref = train_ref.TrainRef(sd, unet_ref.ARCH_S, {k: (e.A, e.B, e.alpha) for k, e in ad.items()}, **kw)
    unet.requires_grad_(False)
    tr = unet.lora_trainer()                      # binds the peft-shaped parameters to the flat device arena
    lora_layers = [p for p in unet.parameters() if p.requires_grad]
    assert len(lora_layers) == 2 * len(tr.slots) and all(p.is_cuda for p in lora_layers)
    opt = torch.optim.AdamW(lora_layers, betas=(0.9, 0.999), eps=1e-8, **kw)
    unet.train()
    p0 = tr.flat_p.cpu().clone()
    for step in range(2):
        lat, noise, t, emb = _batch(2, 32, seed=40 + step)
        noisy = ref.noise_sched.add_noise(lat, noise, t)
        pred = unet(noisy.to(DEV), t.to(DEV), encoder_hidden_states=None, class_labels=emb.to(DEV),
                    cross_attention_kwargs={"scale": 1.0}, return_dict=False)[0]
        loss = F.mse_loss(pred.float(), noise.to(DEV).float(), reduction="mean")
        loss.backward()
        opt.step()
        opt.zero_grad()
        loss_ref = ref.train_step(lat, noise, t, emb)
        assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < 2e-2
    upd, upd_ref = tr.flat_p.cpu() - p0, _flat(ref, tr, "param") - p0
    cos = torch.dot(upd, upd_ref) / (upd.norm() * upd_ref.norm())
    assert cos > 0.98, f"update cosine {cos:.4f}"
    # back to inference: eval + no_grad picks the updated adapters up
    unet.eval()
    lat, _, t, emb = _batch(2, 32, seed=98)
    with torch.no_grad():
        eps = unet(lat.to(DEV), 300, class_labels=emb.to(DEV), return_dict=False)[0]
        lora = unet_ref.LoraSet({p: (A.detach(), B.detach(), al) for p, (A, B, al) in ref.params.items()})
        eps_ref = unet_ref.unet_forward(ref.sd, unet_ref.ARCH_S, lat, 300, emb, lora=lora)
    assert rel(eps, eps_ref) < 3e-2


def test_training_then_validation_pipeline_and_checkpoints_see_the_new_adapters(tmp_path):
    """ADVICE r1 on the device: reference-shaped loop with torch's AdamW (the flat arena changes behind the model's back),
    then the reference's per-epoch validation (`pipe(...)`, train_audioldm_lora.py:599) and checkpoint
    (`get_peft_model_state_dict`, :578) -- twice; both must follow the newest parameters.  Peft init (B = 0): before
    training the adapters have no effect at all, so a stale sync shows up as `pipe == base model`."""
    import torch.nn.functional as F
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.lora import LoraConfig
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=DEV)
    b2.get_peft_model(unet, LoraConfig(r=8, lora_alpha=8, target_modules=["to_q", "to_k", "to_v", "to_out.0"]))
    unet.requires_grad_(False)
    tr = unet.lora_trainer()
    opt = torch.optim.AdamW([p for p in unet.parameters() if p.requires_grad], lr=1e-2)
    pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler())
    x = synthetic.initial_latents(1, 32).to(DEV)
    pos, neg = [t.to(DEV) for t in synthetic.clap_embeddings(1)]
    base = pipe.denoise(x.clone(), pos, neg, 2, 2.5).clone()
    key = next(k for k in b2.get_peft_model_state_dict(unet) if "lora_B" in k)
    outs, sds = [], []
    for rnd_ in range(2):
        unet.train()
        lat, noise, t, emb = _batch(2, 32, seed=80 + rnd_)
        pred = unet(lat.to(DEV), t.to(DEV), encoder_hidden_states=None, class_labels=emb.to(DEV), return_dict=False)[0]
        F.mse_loss(pred.float(), noise.to(DEV).float()).backward()
        opt.step(); opt.zero_grad()
        unet.eval()
        outs.append(pipe.denoise(x.clone(), pos, neg, 2, 2.5).clone())
        sds.append(b2.get_peft_model_state_dict(unet)[key].clone())
        s = tr.slots[key.split("base_model.model.")[1].split(".lora_B")[0]]
        assert torch.equal(sds[-1].reshape(-1), tr.flat_p[s.off_b: s.off_b + s.r * s.c].cpu())
    assert sds[0].abs().max() > 0 and not torch.equal(sds[0], sds[1])
    assert rel(outs[0], base) > 1e-4 and not torch.equal(outs[0], outs[1])


def test_graphed_step_survives_a_lora_scale_change():
    """`train_step_graphed` holds raw pointers into the packed weights.  A `set_lora_scale` round trip between
    steps (validation at another scale) must neither corrupt the next replay (in-place repack keeps pointers) nor be
    ignored; the graphed loss stays equal to the eager loss on the same batch."""
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.train import LoraTrainer
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=DEV)
    unet.load_state_dict(synthetic.random_lora_state_dict(cfg, 8, fmt="peft"), strict=False)
    trainer = LoraTrainer(unet, lr=0.0, weight_decay=0.0)           # lr 0: parameters stay put, losses are comparable
    lat, noise, t, emb = [v.to(DEV) for v in _batch(2, 32, seed=7)]
    l0 = trainer.train_step_graphed(lat, noise, t, emb).item()
    v0, g0 = unet.engine.weights_version, trainer._graph
    unet.engine.set_lora_scale(0.25)
    unet.engine.set_lora_scale(1.0)
    assert unet.engine.weights_version == v0                        # same shapes: overwritten in place
    l1 = trainer.train_step_graphed(lat, noise, t, emb).item()
    assert trainer._graph is g0 and abs(l1 - l0) <= 1e-6 * abs(l0)
    le = trainer.train_step(lat, noise, t, emb).item()
    assert abs(le - l0) < 1e-3 * abs(l0)
    unet.load_state_dict(synthetic.random_lora_state_dict(cfg, 24, fmt="peft"), strict=False)   # r_total 72 > 64: repacked
    assert unet.engine.weights_version != v0
