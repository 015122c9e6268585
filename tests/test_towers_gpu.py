"""The frozen towers (HuBERT, CLIP ViT) through the CUDA plans against the CPU oracle on identical seeded weights.

Tolerance: activations and weights are IEEE fp16 on the tensor cores with fp32 accumulation (the reference itself trains
under ``precision: 16``), the oracle is fp32 — hidden states are compared with max-abs error relative to the tensor's
max-abs value, stated per test.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def _hubert_pair(name):
    from avssl.module import FairseqSpeechEncoder_Hubert
    from oracle import hubert as oh
    from speechclip_b200.init import seeded_init_
    enc = FairseqSpeechEncoder_Hubert(name, feat_select_idx="hidden_states").to(DEV).eval()
    om = seeded_init_(oh.HubertModel(oh.HubertCfg.named(name)), 7122).eval()
    om.load_state_dict(enc.encoder.state_dict())
    return enc, om


@pytest.mark.parametrize("name", ["tiny", "tiny_large"])
def test_hubert_equal_length(name):
    from oracle import hubert as oh
    enc, om = _hubert_pair(name)
    wav = 0.1 * torch.randn(3, 6000, generator=torch.Generator().manual_seed(1))
    states, feat_len = enc(wav.to(DEV))
    with torch.no_grad():
        ref = om.custom_forward(*oh.preprocess_input(list(wav), om.cfg.normalize_wav))["layer_results"]
    assert len(states) == len(ref) == 3 and states[0].shape == ref[0].shape
    for i, (a, b) in enumerate(zip(states, ref)):
        assert rel_err(a.cpu(), b) < 1e-2, (i, rel_err(a.cpu(), b))
    assert feat_len.cpu().tolist() == oh.feat_lengths([6000] * 3, ref[0].shape[1]).tolist()


@pytest.mark.parametrize("name", ["tiny", "tiny_large"])
def test_hubert_variable_length_padding_rules(name):
    """fairseq frame-padding mask (zero before pos-conv, -inf keys) + the wrapper's round(len/320) lengths."""
    from oracle import hubert as oh
    enc, om = _hubert_pair(name)
    g = torch.Generator().manual_seed(2)
    lens = [8000, 3333, 4801, 640]
    wavs = [0.1 * torch.randn(n, generator=g) for n in lens]
    states, feat_len = enc([w.to(DEV) for w in wavs])
    with torch.no_grad():
        out = om.custom_forward(*oh.preprocess_input(wavs, om.cfg.normalize_wav))
    ref, frame_pad = out["layer_results"], out["frame_pad"]
    T = ref[0].shape[1]
    assert feat_len.cpu().tolist() == oh.feat_lengths(lens, T).tolist()
    valid = ~frame_pad
    for i, (a, b) in enumerate(zip(states, ref)):
        a = a.cpu()
        assert rel_err(a[valid], b[valid]) < 1e-2, (i, rel_err(a[valid], b[valid]))
    # the padded 2-D tensor + wav_len entry point gives the same result as the list entry point
    padded = torch.nn.utils.rnn.pad_sequence(wavs, batch_first=True).to(DEV)
    states2, feat_len2 = enc(padded, torch.tensor(lens, device=DEV))
    assert torch.equal(feat_len2, feat_len)
    for a, b in zip(states, states2):
        assert torch.equal(a, b)


def test_hubert_base_full_size_one_utterance():
    """Full-size HuBERT-base architecture, 2 s of audio: conv tap-walk GEMMs at C=512, pos-conv at 16x48, 12 post-LN layers."""
    from oracle import hubert as oh
    enc, om = _hubert_pair("hubert")
    wav = 0.1 * torch.randn(2, 32000, generator=torch.Generator().manual_seed(3))
    states, _ = enc(wav.to(DEV))
    with torch.no_grad():
        collect = {}
        ref = om.custom_forward(wav, None, collect)["layer_results"]
    assert len(states) == 13 and states[0].shape == (2, 99, 768)
    for i, (a, b) in enumerate(zip(states, ref)):
        assert rel_err(a.cpu(), b) < 1.5e-2, (i, rel_err(a.cpu(), b))


def test_hubert_large_full_size_one_utterance():
    """Full-size HuBERT-large structure (LayerNorm extractor, pre-LN, d=1024, 24 layers, 16 heads, pos-conv 16 x 64, waveform
    normalisation), two utterances of different length."""
    from oracle import hubert as oh
    enc, om = _hubert_pair("hubert_large_ll60k")
    g = torch.Generator().manual_seed(6)
    wavs = [0.1 * torch.randn(n, generator=g) for n in (24000, 17000)]
    states, feat_len = enc([w.to(DEV) for w in wavs])
    with torch.no_grad():
        out = om.custom_forward(*oh.preprocess_input(wavs, True))
    ref, valid = out["layer_results"], ~out["frame_pad"]
    assert len(states) == 25 and states[0].shape == (2, 74, 1024)
    for i, (a, b) in enumerate(zip(states, ref)):
        e = rel_err(a.cpu()[valid], b[valid])
        assert e < 2e-2, (i, e)


def test_hubert_weighted_sum_and_training_crop():
    from avssl.module import FairseqSpeechEncoder_Hubert
    enc = FairseqSpeechEncoder_Hubert("tiny", feat_select_idx="weighted_sum", max_audio_len=4000).to(DEV)
    enc.train()
    wav = 0.1 * torch.randn(2, 9000, generator=torch.Generator().manual_seed(4)).to(DEV)
    feat, feat_len = enc(wav, torch.tensor([9000, 2500], device=DEV))
    T = feat.shape[1]
    assert T == 12 and feat_len.tolist() == [12, 8]  # 4000-sample crop -> 12 frames; round(2500/320) = 8
    enc.eval()
    feat_e, _ = enc(wav, torch.tensor([9000, 2500], device=DEV))
    assert feat_e.shape[1] == 27  # eval mode never crops


@pytest.mark.parametrize("name,size", [("tiny", 32), ("ViT-B/32", 224), ("ViT-L/14", 224)])
def test_clip_vit_matches_oracle(name, size):
    from avssl.module import ClipModel
    from oracle import clip as oc
    cm = ClipModel(name).to(DEV).eval()
    om = oc.CLIP(oc.ClipCfg.named(name)).eval()
    om.load_state_dict(cm.model.state_dict())
    img = torch.randn(3, 3, size, size, generator=torch.Generator().manual_seed(5))
    out = cm.encode_image(img.to(DEV))
    with torch.no_grad():
        ref = om.encode_image(img)
    assert out.shape == ref.shape
    assert rel_err(out.cpu(), ref) < 1.5e-2, rel_err(out.cpu(), ref)
    # cosine between matching rows ~ 1: the embeddings that feed the InfoNCE agree in direction
    cos = torch.nn.functional.cosine_similarity(out.cpu(), ref, dim=-1)
    assert cos.min() > 0.9995, cos


def test_no_cpu_fallback():
    from avssl.module import ClipModel, FairseqSpeechEncoder_Hubert
    with pytest.raises(RuntimeError):
        ClipModel("tiny").encode_image(torch.randn(1, 3, 32, 32))
    with pytest.raises(RuntimeError):
        FairseqSpeechEncoder_Hubert("tiny")(torch.randn(1, 4000))
