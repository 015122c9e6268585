"""The oracle against the reference's own outputs (fixtures from tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import speechclip as osc

T = torch.from_numpy


def test_loss_matches_reference(golden):
    for tag in ("small", "mid"):
        z = golden(f"ref_loss_{tag}.npz")
        a, b = T(z["a"]).requires_grad_(), T(z["b"]).requires_grad_()
        ids = T(z["ids"])
        loss = osc.masked_contrastive_loss(a, b, ids, 1 / 0.07)
        loss.backward()
        assert abs(loss.item() - float(z["loss"])) < 1e-6
        assert torch.allclose(a.grad, T(z["da"]), atol=1e-7) and torch.allclose(b.grad, T(z["db"]), atol=1e-7)
        # learnable temperature: multiplier = exp(param)
        tp = T(z["temp_param"]).clone().requires_grad_()
        a2, b2 = T(z["a"]).requires_grad_(), T(z["b"]).requires_grad_()
        lt = osc.masked_contrastive_loss(a2, b2, ids, tp.exp())
        lt.backward()
        assert abs(lt.item() - float(z["loss_t"])) < 1e-6
        assert abs(tp.grad.item() - float(z["dtemp"])) < 1e-5
        assert torch.allclose(a2.grad, T(z["da_t"]), atol=1e-7)
        assert abs(osc.masked_contrastive_loss(T(z["a"]), T(z["b"]), None, 1 / 0.07).item() - float(z["loss_noid"])) < 1e-6


def test_loss_beyond_max_eye():
    # the reference raises IndexError for B > 256 (losses.py:126,211); the oracle generalises eye(B)
    g = torch.Generator().manual_seed(0)
    a = torch.nn.functional.normalize(torch.randn(300, 16, generator=g), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(300, 16, generator=g), dim=-1)
    full = osc.masked_contrastive_loss(a, b, torch.arange(300), 1 / 0.07)
    assert torch.isfinite(full)


def test_weighted_sum_matches_reference(golden):
    z = golden("ref_weighted_sum.npz")
    hidden = list(T(z["hidden"]))
    for norm in (0, 1):
        out = osc.weighted_sum(T(z["weights"]), hidden, bool(norm))
        assert torch.allclose(out, T(z[f"out_norm{norm}"]), atol=1e-6)


def test_keypadding_mask_matches_reference(golden):
    z = golden("ref_keypad.npz")
    assert torch.equal(osc.keypadding_mask(12, T(z["lens"])), T(z["mask"]))


def test_branch_encoder_matches_reference(golden):
    for tag, norm_first in (("postln", False), ("preln", True)):
        z = golden(f"ref_branch_{tag}.npz")
        enc = osc.BranchEncoder(n_layers=1, d_model=64, nhead=8, dim_feedforward=128, norm_first=norm_first).eval()
        enc.load_state_dict({k[3:]: T(z[k]) for k in z.files if k.startswith("sd.")})
        hidden = []
        with torch.no_grad():
            out = enc(T(z["src"]), T(z["kpm"]), hidden)
        valid = ~T(z["kpm"])  # torch's fast path zero-fills padded rows; compare the rows that matter
        assert torch.allclose(out[valid], T(z["out"])[valid], atol=2e-5)
        assert torch.allclose(hidden[0], T(z["h0"]), atol=1e-6)
        assert torch.allclose(hidden[1][valid], T(z["h1"])[valid], atol=2e-5)


def test_retrieval_matches_reference(golden):
    z = golden("ref_retrieval.npz")
    s = T(z["score"])
    ab, ba, mean = osc.mutual_retrieval(s, s.T.contiguous(), T(z["ab"]), T(z["ba"]), [1, 5, 10])
    for i, k in enumerate((1, 5, 10)):
        assert abs(ab[f"recall@{k}"] - z["rAB"][i]) < 1e-4
        assert abs(ba[f"recall@{k}"] - z["rBA"][i]) < 1e-4
        assert abs(mean[f"recall@{k}"] - z["rM"][i]) < 1e-4


def test_scheduler_matches_reference(golden):
    lrs = golden("ref_scheduler.npz")["lrs"]
    mine = [1e-4 * osc.linear_warmup_decay(s, 1e-4, 5, 20, 1e-8) for s in range(20)]
    assert np.allclose(mine, lrs, rtol=1e-12, atol=0)


def test_cascaded_pieces_match_reference(golden):
    z = golden("ref_mha_norm.npz")
    m = osc.AttentionAndNorm(64, 1).eval()
    m.load_state_dict({k[3:]: T(z[k]) for k in z.files if k.startswith("sd.")})
    with torch.no_grad():
        out = m(T(z["src"]), T(z["kpm"]))
    valid = ~T(z["kpm"])
    assert torch.allclose(out[valid], T(z["out"])[valid], atol=2e-6)

    z = golden("ref_kw_bn.npz")
    bn = osc.KwBatchNorm(4, 16, T(z["init_bias"]), T(z["init_scale"])).train()
    x = T(z["x"]).clone().requires_grad_()
    y = bn(x)
    (y * T(z["w"])).sum().backward()
    assert torch.allclose(y, T(z["y_train"]), atol=1e-6) and torch.allclose(x.grad, T(z["dx"]), atol=1e-6)
    assert torch.allclose(bn.bn_layer.weight.grad, T(z["dgamma"]), atol=1e-5) and torch.allclose(bn.bn_layer.bias.grad, T(z["dbeta"]), atol=1e-5)
    assert torch.allclose(bn.bn_layer.running_var, T(z["running_var"]), atol=1e-6)
    bn.eval()
    with torch.no_grad():
        assert torch.allclose(bn(T(z["x"])), T(z["y_eval"]), atol=1e-6)

    z = golden("ref_vq.npz")
    cos = T(z["cos"]).clone().requires_grad_()
    r = osc.simple_vector_quantizer(cos, 0.1, True)
    (r["subword_prob"] * T(z["w"])).sum().backward()
    assert torch.allclose(r["subword_prob"], T(z["subword_prob"]), atol=1e-6) and torch.equal(r["targets"], T(z["targets"]))
    assert torch.allclose(cos.grad, T(z["dcos"]), atol=1e-6)
    for k in ("code_perplexity", "prob_perplexity", "ent_per_t", "diversity_loss"):
        assert torch.allclose(r[k], T(z[k]), atol=1e-5), k
    assert torch.equal(osc.simple_vector_quantizer(T(z["cos"]), 0.1, False)["subword_prob"], T(z["subword_prob_eval"]))


def test_train_mode_dropout_matches_reference_modules(golden):
    """The reference's TransformerEncoder / MultiheadAttentionAndNorm in TRAIN mode (dropout 0.1) with the dropout masks injected
    (make_golden.py): the oracle evaluated with the same masks reproduces every row; with only the row-0 / keyword-row masks (what
    the CUDA path consumes) it reproduces those rows."""
    z = golden("ref_branch_train_dropout.npz")
    enc = osc.BranchEncoder(n_layers=1, d_model=64, nhead=8, dim_feedforward=128).eval()
    enc.load_state_dict({k[3:]: T(z[k]) for k in z.files if k.startswith("sd.")})
    full = dict(attn=T(z["m_attn"]), dropout1=T(z["m_dropout1"]), ffn=T(z["m_ffn"]), dropout2=T(z["m_dropout2"]))
    assert 0.05 < (full["attn"] == 0).float().mean() < 0.15 and abs(full["ffn"].max().item() - 1 / 0.9) < 1e-6
    with torch.no_grad():
        out = enc(T(z["src"]), T(z["kpm"]), masks=full)
        row0 = enc(T(z["src"]), T(z["kpm"]), masks=dict(attn=full["attn"][:, :, 0], dropout1=full["dropout1"][:, 0],
                                                        ffn=full["ffn"][:, 0], dropout2=full["dropout2"][:, 0]))
        plain = enc(T(z["src"]), T(z["kpm"]))
    valid = ~T(z["kpm"])
    assert torch.allclose(out[valid], T(z["out"])[valid], atol=2e-5)
    assert torch.allclose(row0[:, 0], T(z["out"])[:, 0], atol=2e-5)
    assert (plain[:, 0] - T(z["out"])[:, 0]).abs().max() > 0.05   # the masks matter

    z = golden("ref_mha_norm_train_dropout.npz")
    m = osc.AttentionAndNorm(64, 1).eval()
    m.load_state_dict({k[3:]: T(z[k]) for k in z.files if k.startswith("sd.")})
    with torch.no_grad():
        out = m(T(z["src"]), T(z["kpm"]), T(z["m_attn"]))
        rows = m(T(z["src"]), T(z["kpm"]), T(z["m_attn"])[:, :, :3])
    valid = ~T(z["kpm"])
    assert torch.allclose(out[valid], T(z["out"])[valid], atol=2e-5)
    assert torch.allclose(rows[:, :3], T(z["out"])[:, :3], atol=2e-5)
