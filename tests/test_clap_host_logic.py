"""CPU test of the CLAP text tower's host logic: token packing (padded rows dropped), RoBERTa
position ids, fused q / k / v weights, the post-LN layer schedule, pooler and projection of `B200ClapTextEncoder`, with the
kernels replaced by tests/fake_ops.py, against the reference's OWN encoder code: transformers' ClapTextModelWithProjection
in fp32 (/root/reference/script/train/train_audioldm_lora.py:368-369 loads exactly this class; `:513-524` calls it)."""
import pytest
import torch
import torch.nn.functional as F

from audioldm_with_lora_b200.clap import B200ClapTextEncoder, build_text_encoder


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _tokens(nb, L, lens, vocab, g):
    """Right-padded RoBERTa-style ids: <s> ... </s> <pad>*"""
    ids = torch.full((nb, L), 1, dtype=torch.long)
    for i, n in enumerate(lens):
        ids[i, 0] = 0
        ids[i, 1: n - 1] = torch.randint(3, vocab, (max(n - 2, 0),), generator=g)
        ids[i, n - 1] = 2
    return ids, (ids != 1).long()


@pytest.mark.parametrize("lens,L", [([5, 9, 9, 2], 12), ([7], 7), ([3, 3, 3], 16)])
def test_b200_clap_text_encoder_host_logic_matches_transformers(fake_kernels, lens, L):
    enc = build_text_encoder(layers=2, seed=3)
    mine = B200ClapTextEncoder(enc, device="cpu")
    g = torch.Generator().manual_seed(len(lens))
    ids, mask = _tokens(len(lens), L, lens, enc.config.vocab_size, g)
    with torch.no_grad():
        ref = enc(input_ids=ids, attention_mask=mask).text_embeds
    got = mine(ids, attention_mask=mask).text_embeds
    assert got.shape == ref.shape == (len(lens), enc.config.projection_dim) and got.dtype == torch.float32
    assert rel(got, ref) < 2e-2                  # bf16 activations between the layers against fp32
    # what the reference feeds the UNet: the L2-normalised embedding (train_audioldm_lora.py:521-524)
    assert rel(F.normalize(got, dim=-1), F.normalize(ref, dim=-1)) < 2e-2
    cos = (F.normalize(got, dim=-1) * F.normalize(ref, dim=-1)).sum(-1)
    assert float(cos.min()) > 0.999


def test_padding_does_not_change_the_embedding(fake_kernels):
    """The tokenizer pads to 512 (`datasets.py:128-134`): the packed token matrix must make that free AND exact."""
    enc = build_text_encoder(layers=1, seed=4)
    mine = B200ClapTextEncoder(enc, device="cpu")
    g = torch.Generator().manual_seed(0)
    ids, mask = _tokens(2, 6, [6, 4], enc.config.vocab_size, g)
    a = mine(ids, attention_mask=mask).text_embeds
    ids_p = torch.cat([ids, torch.ones(2, 506, dtype=torch.long)], 1)
    mask_p = torch.cat([mask, torch.zeros(2, 506, dtype=torch.long)], 1)
    b = mine(ids_p, attention_mask=mask_p).text_embeds
    assert torch.equal(a, b)


def test_rejects_what_it_cannot_do_exactly(fake_kernels):
    enc = build_text_encoder(layers=1, seed=5)
    mine = B200ClapTextEncoder(enc, device="cpu")
    ids = torch.tensor([[0, 5, 1, 6, 2]])
    with pytest.raises(NotImplementedError):      # a hole in the mask is not right padding
        mine(ids, attention_mask=torch.tensor([[1, 1, 0, 1, 1]]))
    with pytest.raises(ValueError):
        mine(ids, attention_mask=torch.ones(1, 4, dtype=torch.long))


def test_pipeline_encodes_prompts_through_the_text_encoder(fake_kernels):
    """AudioLDMPipeline._encode_prompt with a tokenizer + text encoder (the reference's `pipe("prompt")` path, app.py:14):
    F.normalize(text_embeds), repeated per waveform, "" as the negative prompt."""
    from audioldm_with_lora_b200.pipeline import AudioLDMPipeline
    enc = build_text_encoder(layers=1, seed=6)
    mine = B200ClapTextEncoder(enc, device="cpu")

    class Tok:
        model_max_length = 16

        def __call__(self, texts, padding=None, max_length=None, truncation=None, return_tensors=None):
            rows = []
            for t in texts:
                body = [3 + (ord(ch) % 1000) for ch in t][: max_length - 2]
                rows.append([0] + body + [2] + [1] * (max_length - 2 - len(body)))
            ids = torch.tensor(rows)
            return type("Enc", (), {"input_ids": ids, "attention_mask": (ids != 1).long()})()

    pipe = AudioLDMPipeline.__new__(AudioLDMPipeline)
    pipe.text_encoder, pipe.tokenizer, pipe.device = mine, Tok(), torch.device("cpu")
    pe, ne = pipe._encode_prompt(["rain on a roof", "dog"], 2, True, None, None, None)
    assert pe.shape == (4, 512) and ne.shape == (4, 512)
    assert torch.allclose(pe.norm(dim=-1), torch.ones(4), atol=1e-5) and torch.equal(pe[0], pe[1]) and not torch.equal(pe[1], pe[2])
    with torch.no_grad():
        tok = Tok()(["dog"], max_length=16)
        ref = F.normalize(enc(input_ids=tok.input_ids, attention_mask=tok.attention_mask).text_embeds, dim=-1)
    assert rel(pe[2:3], ref) < 2e-2
