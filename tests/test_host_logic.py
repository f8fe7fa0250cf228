"""CPU tests of the host side: weight packing / K-segment scheduling / LoRA folding / sampler tables /
loader formats / diffusers-shaped API, with the kernels replaced by tests/fake_ops.py (torch-CPU
emulation of the C-ABI semantics), checked against the oracle.  No GPU, no compute through the .so."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import audioldm_with_lora_b200 as b2
from audioldm_with_lora_b200 import _lib, engine as engine_mod, ops, packing, synthetic
from audioldm_with_lora_b200.arch import CONFIGS, UNetConfig, attention_paths, level_sizes
from audioldm_with_lora_b200.lora import LoraConfig, target_linear_paths
from oracle import pipeline_ref, unet_ref
from oracle.ddim_ref import DDIMRef, PNDMRef

ROOT = Path(__file__).resolve().parents[1]
TINY = UNetConfig("tiny", (64, 128, 192, 256))
TINY_SPEC = unet_ref.UNetSpec(block_out_channels=TINY.block_out_channels, time_proj_dim=64)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.fixture(scope="module")
def tiny_weights():
    sd = synthetic.random_unet_state_dict(TINY, seed=0)
    lsd = synthetic.random_lora_state_dict(TINY, 8, fmt="peft")
    ad = b2.parse_lora_state_dict(lsd)
    return sd, lsd, ad, unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()})


# ----------------------------------------------------------------------------- C-ABI surface
def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "b200ldm.h").read_text()
    declared = set(re.findall(r"\b(b200_\w+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    from audioldm_with_lora_b200.build import build_library
    lib = ctypes.CDLL(str(build_library()))
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.b200_version.restype = ctypes.c_int
    assert lib.b200_version() >= 100                       # host-only call; no GPU needed


def test_no_cpu_fallback():
    with pytest.raises(_lib.B200Error):
        _lib.ptr(torch.zeros(4))                           # CPU tensors are refused, not silently computed


# ----------------------------------------------------------------------------- packing / tiling
def test_box_rows_cover_128_pixels():
    for h, w, nb in [(250, 16, 16), (125, 8, 16), (63, 4, 16), (32, 2, 16), (750, 16, 32), (94, 2, 32), (16000, 1, 1)]:
        bh, bni = ops.box_rows(h, w, nb)
        assert bh * w * bni == 128


def test_choose_block_n_valid():
    for n in (8, 128, 256, 384, 640, 768, 1152, 1920, 2048, 5120, 8320):
        for mt in (1, 8, 32, 125, 512):
            for geglu in (False, True):
                bn = ops.choose_block_n(n, mt, geglu)
                assert 64 <= bn <= 256 and bn % (128 if geglu else 64) == 0


def test_conv_weight_k_order_and_geglu_interleave():
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).view(2, 3, 3, 3)
    k = packing.conv3x3_to_k(w)                            # [2, 9 * 64], tap-major
    assert k.shape == (2, 576)
    assert k[1, (1 * 3 + 2) * 64 + 2] == w[1, 2, 1, 2] and k[:, 3:64].abs().sum() == 0
    wf = torch.randn(512, 64)
    pw = packing.pack([wf], torch.arange(512.0), 128, 1, 64, geglu=True, device="cpu")
    assert pw.n_valid == 256 and pw.n_pad == 512
    # tile 1 = [values 64..127 | gates 64..127]
    assert torch.equal(pw.w[128:192].float(), wf[64:128].to(torch.bfloat16).float())
    assert torch.equal(pw.w[192:256].float(), wf[256 + 64:256 + 128].to(torch.bfloat16).float())
    assert pw.bias[192].item() == 256 + 64


def test_arena_is_deterministic_and_leak_free():
    ar = engine_mod.Arena(1 << 20, "cpu")
    a = ar.alloc((100, 64), torch.bfloat16); b = ar.alloc((10,), torch.float32)
    pa, pb = a.data_ptr(), b.data_ptr()
    ar.release(a); ar.release(b)
    assert ar.free == [(0, 1 << 20)] and not ar.live
    assert ar.alloc((100, 64), torch.bfloat16).data_ptr() == pa and ar.alloc((10,), torch.float32).data_ptr() == pb
    with pytest.raises(MemoryError):
        ar.alloc((1 << 21,), torch.uint8)


# ----------------------------------------------------------------------------- engine vs oracle
@pytest.mark.parametrize("h", [25, 16])
def test_engine_matches_oracle_per_layer(fake_kernels, tiny_weights, h):
    sd, _, ad, ora = tiny_weights
    eng = engine_mod.UNetEngine(TINY, sd, "cpu")
    eng.set_lora(ad, 1.0)
    x = synthetic.initial_latents(2, h)
    pos, _ = synthetic.clap_embeddings(2)
    to, te = {}, {}
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, TINY_SPEC, x, 501, pos, lora=ora, taps=to)
        out = eng.forward(x, 501, pos, taps=te)
    assert rel(out, ref) < 3e-2                            # bf16 activations, fp32 accumulation
    assert rel(te["emb"], to["emb"]) < 1e-5
    for k in to:
        if k in te and k != "emb":
            assert rel(te[k], to[k]) < 3e-2, k


def test_engine_lora_scale_and_removal(fake_kernels, tiny_weights):
    sd, _, ad, _ = tiny_weights
    eng = engine_mod.UNetEngine(TINY, sd, "cpu")
    x = synthetic.initial_latents(1, 16); pos, _ = synthetic.clap_embeddings(1)
    base = eng.forward(x, 300, pos)
    eng.set_lora(ad, 1.0)
    with_lora = eng.forward(x, 300, pos)
    eng.set_lora_scale(0.0)
    off = eng.forward(x, 300, pos)
    eng.set_lora(None)
    assert rel(with_lora, base) > 1e-3 and rel(off, base) < 1e-6 and rel(eng.forward(x, 300, pos), base) == 0.0
    with torch.no_grad():
        ref2 = unet_ref.unet_forward(sd, TINY_SPEC, x, 300, pos,
                                     lora=unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()}, scale=2.0))
    eng.set_lora(ad, 2.0)
    assert rel(eng.forward(x, 300, pos), ref2) < 3e-2


def test_reference_lora_config_q_v_only(fake_kernels, tiny_weights):
    """The reference trains r=2 on to_q,to_v only (train_audioldm_lora.py:378-383)."""
    sd = tiny_weights[0]
    lsd = synthetic.random_lora_state_dict(TINY, 2, targets=("to_q", "to_v"), fmt="peft_sd")
    ad = b2.parse_lora_state_dict(lsd, alpha=2)
    assert len(ad) == 64
    eng = engine_mod.UNetEngine(TINY, sd, "cpu"); eng.set_lora(ad)
    x = synthetic.initial_latents(1, 16); pos, _ = synthetic.clap_embeddings(1)
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, TINY_SPEC, x, 10, pos, lora=unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()}))
    assert rel(eng.forward(x, 10, pos), ref) < 3e-2


# ----------------------------------------------------------------------------- sampler tables / pipeline loop
@pytest.mark.parametrize("n", [10, 50, 200])
def test_ddim_table_matches_scheduler_step(n):
    s = b2.DDIMScheduler(); s.set_timesteps(n)
    o = DDIMRef(); assert o.set_timesteps(n).tolist() == s.unet_timesteps()
    tab = s.step_table()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 6, 16, generator=g)
    for i, t in enumerate(s.unet_timesteps()):
        e = torch.randn(2, 8, 6, 16, generator=g)
        ref = o.step(e, t, x)
        x = tab[i, 0] * x + tab[i, 1] * e
        assert rel(x, ref) < 2e-6
        x = ref


def test_pndm_table_matches_plms(fake_kernels):
    s = b2.PNDMScheduler(); s.set_timesteps(10)
    o = PNDMRef(); assert o.set_timesteps(10).tolist() == s.unet_timesteps()
    tab = s.step_table()
    nb, hw, c = 1, 32, 8
    g = torch.Generator().manual_seed(0)
    x = torch.randn(nb, hw, c, generator=g)
    xo = x.clone()
    xs, hist, step = torch.zeros_like(x), torch.zeros(4, nb, hw, c), torch.zeros(1, dtype=torch.int32)
    for i, t in enumerate(s.unet_timesteps()):
        e = torch.randn(nb, hw, c, generator=g)
        xo = o.step(e, t, xo)
        fake_kernels.sampler_step(e, x, xs, hist, tab, step, 1.0, False, nb, hw, c, 64, None)
        assert rel(x, xo) < 1e-5, i
    assert int(step) == 11


@pytest.mark.parametrize("sched", ["ddim", "pndm"])
def test_pipeline_loop_matches_oracle(fake_kernels, tiny_weights, sched):
    sd, lsd, ad, ora = tiny_weights
    unet = b2.UNet2DConditionModel(TINY, sd, device="cpu")
    unet.load_state_dict(lsd, strict=False)
    pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler() if sched == "ddim" else b2.PNDMScheduler(), use_cuda_graph=False)
    x = synthetic.initial_latents(2, 25); pos, neg = synthetic.clap_embeddings(2)
    tr, tro = [], []
    out = pipe.denoise(x.clone(), pos, neg, 6, 2.5, trace=tr)
    with torch.no_grad():
        if sched == "ddim":
            ref = pipeline_ref.denoise_loop(sd, TINY_SPEC, pos, neg, x.clone(), 6, 2.5, lora=ora, trace=tro)
        else:
            s = PNDMRef(); lat = x.clone(); emb = torch.cat([neg, pos])
            for t in s.set_timesteps(6):
                e = unet_ref.unet_forward(sd, TINY_SPEC, torch.cat([lat] * 2), t, emb, lora=ora)
                eu, et = e.chunk(2)
                lat = s.step(eu + 2.5 * (et - eu), int(t), lat); tro.append(lat.clone())
            ref = lat
    assert len(tr) == len(tro)
    for a, b in zip(tr, tro):
        assert rel(a, b) < 2e-2                            # per-step latent rel-L2, bf16 bound of BASELINE.json
    assert rel(out, ref) < 2e-2
    # no guidance: single conditional batch
    out1 = pipe.denoise(x.clone(), pos, None, 3, 1.0)
    with torch.no_grad():
        ref1 = pipeline_ref.denoise_loop(sd, TINY_SPEC, pos, neg, x.clone(), 3, 1.0, lora=ora) if sched == "ddim" else None
    if ref1 is not None:
        assert rel(out1, ref1) < 2e-2


# ----------------------------------------------------------------------------- LoRA loader formats
def test_three_key_formats_load_identically(tiny_weights):
    ad0 = tiny_weights[2]
    for fmt in ("peft", "peft_sd", "diffusers"):
        ad = b2.parse_lora_state_dict(synthetic.random_lora_state_dict(TINY, 8, fmt=fmt))
        assert set(ad) == set(ad0)
        for k in ad:
            assert torch.equal(ad[k].A, ad0[k].A) and torch.equal(ad[k].B, ad0[k].B)
    # diffusers format with the `unet.` prefix (load_attn_procs files) and a full accelerate state (base tensors ignored)
    dsd = {"unet." + k: v for k, v in synthetic.random_lora_state_dict(TINY, 8, fmt="diffusers").items()}
    assert set(b2.parse_lora_state_dict(dsd)) == set(ad0)
    full = dict(synthetic.random_lora_state_dict(TINY, 8, fmt="peft"))
    full["base_model.model.conv_in.weight"] = torch.zeros(1)
    full["base_model.model." + attention_paths(TINY)[0] + ".to_q.base_layer.weight"] = torch.zeros(1)
    assert set(b2.parse_lora_state_dict(full)) == set(ad0)
    # peft -> diffusers conversion (train_audioldm_lora.py:578) round-trips
    conv = b2.convert_state_dict_to_diffusers(b2.to_peft_state_dict(ad0))
    assert all(".lora.down.weight" in k or ".lora.up.weight" in k for k in conv)
    assert set(b2.parse_lora_state_dict(conv)) == set(ad0)


def test_loader_errors_and_target_rule():
    sd = synthetic.random_lora_state_dict(TINY, 4, fmt="peft")
    k = next(k for k in sd if "lora_B" in k)
    bad = dict(sd); bad.pop(k)
    with pytest.raises(KeyError):
        b2.parse_lora_state_dict(bad)
    assert len(target_linear_paths(TINY, ("to_q", "to_v"))) == 64
    assert len(target_linear_paths(TINY, ("to_q", "to_k", "to_v", "to_out.0"))) == 128
    with pytest.raises(NotImplementedError):
        LoraConfig(lora_dropout=0.1)


def test_safetensors_file_roundtrip(tmp_path, fake_kernels, tiny_weights):
    from safetensors.torch import save_file
    sd, lsd, ad, _ = tiny_weights
    save_file({k: v.contiguous() for k, v in synthetic.random_lora_state_dict(TINY, 8, fmt="diffusers").items()},
              str(tmp_path / "pytorch_lora_weights.safetensors"))
    unet = b2.UNet2DConditionModel(TINY, sd, device="cpu")
    unet.load_attn_procs(str(tmp_path))                    # app.py:11
    assert set(unet.engine.lora) == set(ad)
    got = b2.get_peft_model_state_dict(unet)
    assert all(k.startswith("base_model.model.") and ".lora_" in k for k in got) and len(got) == 2 * len(ad)


# ----------------------------------------------------------------------------- diffusers-shaped model API
class TorchSDPAProcessor:
    """Plain torch processor with the AttnProcessor2_0 call signature (what the reference runs)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **kw):
        import torch.nn.functional as F
        x = hidden_states.float()
        b, s, c = x.shape

        def lin(m, t):
            y = F.linear(t, m.weight.float(), None if m.bias is None else m.bias.float())
            if hasattr(m, "lora_A") and len(m.lora_A):
                y = y + m.lora_B["default"](m.lora_A["default"](t)) * m.scaling["default"]
            return y
        q, k, v = (lin(getattr(attn, n), x).view(b, s, attn.heads, -1).transpose(1, 2) for n in ("to_q", "to_k", "to_v"))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, s, c)
        return attn.to_out[1](lin(attn.to_out[0], o)).to(hidden_states.dtype)


def test_attn_processor_api(fake_kernels, tiny_weights):
    sd, lsd, ad, ora = tiny_weights
    unet = b2.UNet2DConditionModel(TINY, sd, device="cpu")
    procs = unet.attn_processors
    assert len(procs) == 32 and "down_blocks.1.attentions.0.transformer_blocks.0.attn1.processor" in procs
    assert "mid_block.attentions.0.transformer_blocks.0.attn2.processor" in procs
    assert all(isinstance(p, b2.B200AttnProcessor) for p in procs.values())
    with pytest.raises(ValueError, match="number of attention layers"):
        unet.set_attn_processor({"x.processor": TorchSDPAProcessor()})
    b2.get_peft_model(unet, LoraConfig(r=8, lora_alpha=8))            # B = 0 at init
    unet.load_state_dict(lsd, strict=False)
    x = synthetic.initial_latents(1, 16); pos, _ = synthetic.clap_embeddings(1)
    native = unet(x, 400, class_labels=pos, cross_attention_kwargs={"scale": 1.0}, return_dict=False)[0]
    # foreign processors through the seam: every attention module, then only one
    unet.set_attn_processor(TorchSDPAProcessor())
    foreign = unet(x, 400, class_labels=pos).sample
    assert rel(foreign, native) < 2e-2                     # fp32 foreign processor vs bf16 native attention
    unet.set_default_attn_processor()
    d = unet.attn_processors
    d["up_blocks.1.attentions.2.transformer_blocks.0.attn2.processor"] = TorchSDPAProcessor()
    unet.set_attn_processor(d)
    assert rel(unet(x, 400, class_labels=pos).sample, native) < 1e-2
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, TINY_SPEC, x, 400, pos, lora=ora)
    assert rel(native, ref) < 3e-2
    with pytest.raises(NotImplementedError):
        unet(x, 400, encoder_hidden_states=torch.zeros(1, 1, 8), class_labels=pos)
    with pytest.raises(ValueError):
        unet(x, 400)
    with pytest.raises(ValueError, match="already exists"):
        unet.add_adapter(LoraConfig(r=8, lora_alpha=8))


def test_peft_init_gaussian_b_zero(fake_kernels, tiny_weights):
    sd = tiny_weights[0]
    unet = b2.UNet2DConditionModel(TINY, sd, device="cpu")
    x = synthetic.initial_latents(1, 16); pos, _ = synthetic.clap_embeddings(1)
    base = unet(x, 5, class_labels=pos).sample
    unet.add_adapter(LoraConfig(r=2, lora_alpha=2, target_modules=["to_q", "to_v"]))
    assert len(unet.engine.lora) == 64
    e = next(iter(unet.engine.lora.values()))
    assert e.B.abs().sum() == 0 and 0.2 < e.A.std() < 0.8              # N(0, (1/r)^2), r = 2
    assert rel(unet(x, 5, class_labels=pos).sample, base) < 1e-6      # B = 0 -> identical


# ----------------------------------------------------------------------------- pipeline __call__ surface
def test_pipeline_call_contract(fake_kernels, tiny_weights):
    from audioldm_with_lora_b200 import tail
    sd = tiny_weights[0]
    unet = b2.UNet2DConditionModel(TINY, sd, device="cpu")
    pipe = b2.AudioLDMPipeline(unet, vae=tail.random_vae_decoder(7), vocoder=tail.build_vocoder(0),
                               tail_dtype=torch.float32, use_cuda_graph=False)
    pos, neg = synthetic.clap_embeddings(2)
    with pytest.raises(ValueError, match="prompt_embeds"):
        pipe("a hip hop beat")                                         # no text encoder offline
    with pytest.raises(ValueError):
        pipe(prompt_embeds=pos, negative_prompt_embeds=neg[:1], audio_length_in_s=0.64)
    with pytest.raises(ValueError):
        pipe(prompt_embeds=pos, negative_prompt_embeds=neg, audio_length_in_s=0.01)
    seen = []
    out = pipe(prompt_embeds=pos, negative_prompt_embeds=neg, audio_length_in_s=0.64, num_inference_steps=3,
               generator=torch.Generator().manual_seed(0), callback=lambda i, t, l: seen.append((i, t, tuple(l.shape))))
    assert out.audios.dtype == np.float32 and out.audios.shape == (2, int(0.64 * 16000))
    assert seen == [(0, 667, (2, 8, 16, 16)), (1, 334, (2, 8, 16, 16)), (2, 1, (2, 8, 16, 16))]
    lat = pipe(prompt_embeds=pos, negative_prompt_embeds=neg, audio_length_in_s=0.64, num_inference_steps=2,
               num_waveforms_per_prompt=2, generator=torch.Generator().manual_seed(0), output_type="latent",
               return_dict=False)[0]
    assert tuple(lat.shape) == (4, 8, 16, 16)
    # default length = sample_size * 4 * 0.01 s = 5.12 s -> latent height 128
    assert level_sizes(128)[3] == (16, 2)


def test_branched_middle_region_equals_single_chain(fake_kernels):
    """engine.forward_nhwc(mid_branches=k): the deep levels run as k sub-batch chains (parallel streams on the GPU,
    sequential here) -- same result as the single chain, no arena leak, region boundaries as documented."""
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.arch import UNetConfig
    from audioldm_with_lora_b200.engine import LATENT_C_PAD
    tiny = UNetConfig("tiny", (64, 128, 192, 256))
    unet = b2.UNet2DConditionModel(tiny, synthetic.random_unet_state_dict(tiny, seed=0), device="cpu")
    unet.load_state_dict(synthetic.random_lora_state_dict(tiny, 4, fmt="peft"), strict=False)
    eng = unet.engine
    prog = eng.program()
    i0, i1, need = eng.middle_region()
    assert prog[i0][0] == "down" and prog[i0][2] == 1 and prog[i1 - 1][0] == "up" and prog[i1 - 1][2] == 2 and need == 0
    assert all(st[2] >= 2 for st in prog[i0:i1] if st[0] in ("res", "tfm"))
    assert all(st[2] < 2 for st in prog[:i0] + prog[i1:] if st[0] in ("res", "tfm"))
    nb, h, w = 4, 24, 16
    xin = torch.zeros(nb, h * w, LATENT_C_PAD, dtype=torch.bfloat16)
    xin[:, :, :8] = torch.randn(nb, h * w, 8, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    silu = torch.randn(nb, tiny.temb_channels, generator=torch.Generator().manual_seed(4)).to(torch.bfloat16)
    outs = []
    for k in (1, 2, 4, 3):                       # 3 does not divide 4 -> falls back to 2
        eps = torch.zeros(nb, h * w, 8)
        eng.forward_nhwc(xin, silu, nb, h, w, eps, mid_branches=k)
        assert not eng.arena.live
        outs.append(eps)
    for o in outs[1:]:
        # not bit-equal: a different batch size changes the fp32 summation order inside the matmuls, and the
        # bf16 roundings between layers amplify it to the usual bf16 noise floor (a logic error would be O(1))
        assert ((o - outs[0]).norm() / outs[0].norm()).item() < 2e-2
