"""The oracle's restatement of the ABSENT third-party towers against an independent implementation.

fairseq (HuBERT) and openai/CLIP cannot be installed here, so the tower oracles are unpinned against
them; transformers' Hubert / CLIP implement the same published architectures, and must agree with
the oracle on identical weights (equal-length inputs: the padding rules legitimately differ).
"""
import numpy as np
import pytest
import torch

from oracle import clip as oc
from oracle import hubert as oh
from speechclip_b200.init import seeded_init_

T = torch.from_numpy


@pytest.mark.parametrize("name", ["tiny", "tiny_large"])
def test_hubert_oracle_vs_hf_fixture(golden, name):
    z = golden(f"hf_hubert_{name}.npz")
    om = seeded_init_(oh.HubertModel(oh.HubertCfg.named(name)), 7122).eval()
    with torch.no_grad():
        out = om.custom_forward(T(z["wav"]), None)
    hs = out["layer_results"]
    n = len(hs)
    for i in range(n - 1):
        assert torch.allclose(hs[i], T(z[f"h{i}"]), atol=2e-5), i
    last = hs[-1] if not om.cfg.layer_norm_first else om.encoder.layer_norm(hs[-1])
    assert torch.allclose(last, T(z[f"h{n - 1}"]), atol=2e-5)


def test_hubert_oracle_vs_hf_live_base_shape():
    """Full-size base architecture, live against transformers (one short utterance)."""
    from tests.hf_map import hf_hubert_from_oracle
    om = seeded_init_(oh.HubertModel(oh.HubertCfg.named("hubert")), 7122).eval()
    hf = hf_hubert_from_oracle(om)
    wav = 0.1 * torch.randn(1, 16000, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        mine = om.custom_forward(wav, None)["layer_results"]
        theirs = hf(wav, output_hidden_states=True).hidden_states
    assert len(mine) == len(theirs) == 13 and mine[0].shape == (1, 49, 768)
    for a, b in zip(mine, theirs):
        assert torch.allclose(a, b, atol=5e-4), (a - b).abs().max()


def test_frame_rules_match_survey_probe():
    # len 48000 -> conv 149 / round 150 / fairseq-valid 150 ; len 48160 -> 150 / 150 / 151
    for n, conv, rnd, valid in ((48000, 149, 150, 150), (48160, 150, 150, 151)):
        assert oh.conv_out_length(n) == conv
        assert oh.feat_lengths([n], 319).item() == rnd
        pad = ~(torch.arange(102400)[None] < torch.tensor([[n]]))
        assert int((~oh.HubertModel.frame_padding_mask(319, pad)).sum()) == valid
    assert oh.conv_out_length(102400) == 319 and oh.conv_out_length(16000) == 49


def test_clip_oracle_vs_hf_fixture(golden):
    z = golden("hf_clip_tiny.npz")
    om = seeded_init_(oc.CLIP(oc.ClipCfg.named("tiny")), 7122).eval()
    with torch.no_grad():
        assert torch.allclose(om.encode_image(T(z["img"])), T(z["image_embeds"]), atol=2e-5)
        assert torch.allclose(om.encode_text(T(z["tok"])), T(z["text_embeds"]), atol=2e-5)


def test_clip_oracle_vs_hf_live_vitb32():
    from tests.hf_map import hf_clip_from_oracle
    om = seeded_init_(oc.CLIP(oc.ClipCfg.named("ViT-B/32")), 7122).eval()
    hv, _ = hf_clip_from_oracle(om)
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        a, b = om.encode_image(img), hv(pixel_values=img).image_embeds
    assert a.shape == (2, 512) and torch.allclose(a, b, atol=5e-4), (a - b).abs().max()


@pytest.mark.parametrize("name,atol", [("hubert", 2e-5), ("hubert_large_ll60k", 1e-4)])
def test_hubert_oracle_vs_torchaudio_through_its_fairseq_key_map(name, atol):
    """torchaudio's wav2vec2 / HuBERT is a re-implementation made to load fairseq checkpoints: `import_fairseq._convert_state_dict`
    is ITS map from fairseq's state-dict keys to its own.  The oracle's state dict (fairseq key names, as the reference's
    checkpoints have them: speech_encoder_plus.py:387,499-504) must pass through that map with nothing left over, load into
    torchaudio's model of the same architecture, and give the same hidden states — base (GroupNorm extractor, post-LN) and
    large (LayerNorm extractor, pre-LN, conv_bias False; the states are collected before the encoder's final LayerNorm)."""
    torchaudio = pytest.importorskip("torchaudio")
    from torchaudio.models.wav2vec2.utils.import_fairseq import _map_key
    om = seeded_init_(oh.HubertModel(oh.HubertCfg.named(name)), 7122).eval()
    mapped = {}
    for k, v in om.state_dict().items():
        nk = _map_key(k)  # raises ValueError on a key fairseq's HubertModel would not have
        if nk is not None:
            mapped[nk] = v
    ta = (torchaudio.models.hubert_base() if name == "hubert" else torchaudio.models.hubert_large()).eval()
    missing, unexpected = ta.load_state_dict(mapped, strict=False)
    assert not missing and set(unexpected) <= {"label_embs_concat"}, (missing, unexpected)
    wav = (0.1 if name == "hubert" else 1.0) * torch.randn(1, 8000, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        mine = om.custom_forward(wav, None)["layer_results"]
        theirs, _ = ta.extract_features(wav)
    assert len(mine) == len(theirs) + 1  # the oracle also returns the encoder input (hidden state 0)
    for a, b in zip(mine[1:], theirs):
        assert torch.allclose(a, b, atol=atol), (a - b).abs().max()
