"""GPU parity of the CLAP text tower on the sm_100a kernels: `b200_embed_layernorm` and the GELU /
ReLU / tanh linear epilogues against torch, and `B200ClapTextEncoder` against the reference's own encoder code (transformers
ClapTextModelWithProjection, fp32; /root/reference/script/train/train_audioldm_lora.py:368-369, :513-524) on the same weights."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
KTOL = 5e-3         # single kernel vs torch fp32 on the same inputs (bf16 output rounding)
TOL = 2e-2          # whole encoder: bf16 activations through 12 post-LN layers against fp32


@pytest.fixture(scope="module", autouse=True)
def _need_cuda_and_lib():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the hot path has no CPU fallback")
    from audioldm_with_lora_b200 import _lib
    _lib.load()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("m,c,vocab,npos", [(37, 768, 1000, 514), (1, 768, 50265, 514), (300, 1024, 777, 40)])
def test_embed_layernorm_matches_torch(m, c, vocab, npos):
    from audioldm_with_lora_b200 import ops
    g = torch.Generator().manual_seed(m)
    word, pos = torch.randn(vocab, c, generator=g), torch.randn(npos, c, generator=g)
    type0, gamma, beta = torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    ids = torch.randint(0, vocab, (m,), generator=g, dtype=torch.int32)
    pids = torch.randint(0, npos, (m,), generator=g, dtype=torch.int32)
    ref = F.layer_norm((word[ids.long()] + type0) + pos[pids.long()], (c,), gamma, beta, 1e-12)
    y = torch.full((m, c), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.embed_layernorm(ids.to(DEV), pids.to(DEV), word.to(DEV), pos.to(DEV), type0.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-12, y)
    assert torch.isfinite(y.float()).all() and rel(y, ref) < KTOL


@pytest.mark.parametrize("act", ["gelu", "relu", "tanh", "none"])
@pytest.mark.parametrize("m,k,n", [(37, 768, 3072), (5, 3072, 768), (300, 768, 512)])
def test_linear_epilogue_activations_match_torch(act, m, k, n):
    from audioldm_with_lora_b200 import ops, packing
    g = torch.Generator().manual_seed(m + n)
    w, b = torch.randn(n, k, generator=g) / k ** 0.5, torch.randn(n, generator=g) * 0.3
    x = (torch.randn(m, k, generator=g) * 1.5).to(torch.bfloat16)
    res = torch.randn(m, n, generator=g).to(torch.bfloat16)
    pw = packing.pack([w], b, ops.choose_tiling(n, 1, k // 64, allow_split=False)[0], 1, k, device=DEV)
    z = x.float() @ w.to(torch.bfloat16).float().T + b + res.float()
    ref = {"gelu": F.gelu, "relu": F.relu, "tanh": torch.tanh, "none": lambda t: t}[act](z)
    out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=DEV)
    kw = {"gelu": dict(act_tanh=2), "relu": dict(act_slope=0.0), "tanh": dict(act_tanh=True), "none": {}}[act]
    ops.conv1d(pw, x.to(DEV), 1, m, out, dh0=0, dh_step=1, residual=res.to(DEV), **kw)
    assert torch.isfinite(out.float()).all() and rel(out, ref) < KTOL


@pytest.mark.parametrize("lens,L,layers", [([5, 9, 9, 2], 12, 12), ([77], 77, 2), ([3] * 8 + [130, 1], 512, 2)])
def test_b200_clap_text_encoder_matches_transformers(lens, L, layers):
    from audioldm_with_lora_b200 import _lib
    from audioldm_with_lora_b200.clap import B200ClapTextEncoder, build_text_encoder
    enc = build_text_encoder(layers=layers, seed=7)
    mine = B200ClapTextEncoder(enc, device=DEV)
    g = torch.Generator().manual_seed(len(lens))
    ids = torch.full((len(lens), L), 1, dtype=torch.long)
    for i, n in enumerate(lens):
        ids[i, :n] = torch.randint(3, enc.config.vocab_size, (n,), generator=g)
        ids[i, 0] = 0
    mask = (torch.arange(L)[None] < torch.tensor(lens)[:, None]).long()
    with torch.no_grad():
        ref = enc(input_ids=ids, attention_mask=mask).text_embeds
    n0 = _lib.launch_count
    got = mine(ids, attention_mask=mask).text_embeds
    assert _lib.launch_count - n0 >= 1 + layers * 7 + 3          # every layer ran on the C-ABI kernels
    assert got.is_cuda and got.shape == ref.shape and torch.isfinite(got).all()
    assert rel(got, ref) < TOL
    cos = (F.normalize(got.cpu(), dim=-1) * F.normalize(ref, dim=-1)).sum(-1)
    assert float(cos.min()) > 0.999
    # the pipeline re-hosts the torch module and feeds the UNet F.normalize(text_embeds)
    from audioldm_with_lora_b200.pipeline import AudioLDMPipeline
    pipe = AudioLDMPipeline.__new__(AudioLDMPipeline)
    pipe.text_encoder, pipe.device = mine, torch.device(DEV)
    pipe.tokenizer = type("Tok", (), {"model_max_length": L, "__call__": lambda self, texts, **kw: type("E", (), {
        "input_ids": ids[: len(texts)], "attention_mask": mask[: len(texts)]})()})()
    pe, ne = pipe._encode_prompt(["a"] * len(lens), 1, False, None, None, None)
    assert ne is None and rel(pe, F.normalize(ref, dim=-1)) < TOL
