"""Train-mode dropout of the trainable branches (reference: TransformerModels.py:55-75 ->
nn.TransformerEncoderLayer(dropout=0.1) for the parallel branch, :110-117 -> nn.MultiheadAttention(dropout=0.1) for the cascaded
one; spchclp_p.yaml:27).  The CUDA path draws its masks from a counter-based RNG (scb_dropout_mask) and regenerates them in the
backward pass; the test materialises the SAME masks and hands them to the CPU oracle, whose train-mode arithmetic is pinned
against the reference's own modules (tests/golden/ref_*_train_dropout.npz)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-8)).item()


def test_dropout_mask_generator_statistics_and_determinism():
    from speechclip_b200 import ops
    st = torch.tensor([1234567, 5], dtype=torch.int64, device=DEV)
    n = 1 << 20
    m = ops.dropout_mask(st, 0, 0.1, n)
    vals = torch.unique(m)
    assert vals.numel() == 2 and vals[0] == 0 and abs(vals[1].item() - 1 / 0.9) < 1e-6
    drop = (m == 0).float().mean().item()
    assert abs(drop - 0.1) < 2e-3, drop                           # 5 sigma = 1.5e-3
    assert torch.equal(m, ops.dropout_mask(st, 0, 0.1, n))        # pure function of (seed, step, site, index)
    assert abs((m[:-1] * m[1:]).mean().item() - 1.0) < 5e-3       # neighbours uncorrelated: E[m_i m_{i+1}] = 1
    for other in (torch.tensor([1234567, 6], dtype=torch.int64, device=DEV), torch.tensor([1234568, 5], dtype=torch.int64, device=DEV)):
        agree = ((ops.dropout_mask(other, 0, 0.1, n) == 0) == (m == 0)).float().mean().item()
        assert abs(agree - 0.82) < 5e-3, agree                    # independent draws: 0.9^2 + 0.1^2
    agree = ((ops.dropout_mask(st, 1, 0.1, n) == 0) == (m == 0)).float().mean().item()
    assert abs(agree - 0.82) < 5e-3, agree
    st2 = st.clone()
    ops.rng_advance(st2)
    assert st2.tolist() == [1234567, 6]
    x = torch.randn(1000, device=DEV)
    y = ops.dropout_rows(x, torch.empty_like(x), (0.25, st, 3))
    assert torch.equal(y, x * ops.dropout_mask(st, 3, 0.25, 1000))
    assert ((ops.dropout_mask(st, 0, 0.5, n) == 0).float().mean().item() - 0.5) < 3e-3


def _parallel_pair(d=64, heads=4, ffn=128, out=32, p=0.1):
    from avssl.base import OrderedNamespace
    from avssl.model.kwClip import KW_ParallelBranch
    from oracle import speechclip as osc
    from speechclip_b200.configs import parallel_config
    cfg = parallel_config("tiny")
    cfg["model_settings"]["parallel_branch"]["transformer_args"].update(d_model=d, nhead=heads, dim_feedforward=ffn, dropout=p)
    torch.manual_seed(0)
    mine = KW_ParallelBranch(OrderedNamespace(cfg), d, out)
    with torch.no_grad():
        for n, prm in mine.named_parameters():
            if n.endswith("bias") or "norm" in n:
                prm.add_(0.05 * torch.randn(prm.shape))
    ref = osc.ParallelBranch(d, out, n_layers=1, nhead=heads, dim_feedforward=ffn).eval()
    ref.load_state_dict(mine.state_dict())
    return mine.to(DEV), ref


@pytest.mark.parametrize("d,heads,ffn,B,T", [(64, 4, 128, 6, 21), (768, 8, 3072, 5, 319)])
def test_parallel_branch_train_mode_dropout_vs_oracle_with_the_same_masks(d, heads, ffn, B, T):
    from speechclip_b200 import ops
    from speechclip_b200.head import SITE_ATTN, SITE_DROPOUT1, SITE_DROPOUT2, SITE_FFN
    p = 0.1
    mine, ref = _parallel_pair(d, heads, ffn, 32, p)
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(B, T, d, generator=g).half().float()
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[0] = T
    w = torch.randn(B, 32, generator=g)
    mine.train()
    fd = feat.to(DEV).requires_grad_(True)
    out = mine(fd, lens.to(DEV))
    state = mine._scb_dropout.get(DEV).clone()      # the state the forward used (advanced once, before drawing)
    assert state[1].item() == 1
    (out * w.to(DEV)).sum().backward()
    Tk = T + 1
    masks = dict(attn=ops.dropout_mask(state, SITE_ATTN, p, B * heads * Tk).view(B, heads, Tk).cpu(),
                 dropout1=ops.dropout_mask(state, SITE_DROPOUT1, p, B * d).view(B, d).cpu(),
                 ffn=ops.dropout_mask(state, SITE_FFN, p, B * ffn).view(B, ffn).cpu(),
                 dropout2=ops.dropout_mask(state, SITE_DROPOUT2, p, B * d).view(B, d).cpu())
    fr = feat.clone().requires_grad_(True)
    oout = ref(fr, lens, masks=masks)
    (oout * w).sum().backward()
    assert rel_err(out.detach().cpu(), oout.detach()) < 5e-3
    # the masks matter: the eval-mode result is far away
    with torch.no_grad():
        assert rel_err(ref(feat, lens), oout.detach()) > 5e-2
    oparams = dict(ref.named_parameters())
    for name, prm in mine.named_parameters():
        e = rel_err(prm.grad.cpu(), oparams[name].grad)
        assert e < 3e-2, (name, e)
    assert rel_err(fd.grad.cpu(), fr.grad) < 3e-2
    # a second step draws new masks; eval mode draws none and is deterministic
    out2 = mine(fd, lens.to(DEV))
    assert mine._scb_dropout.get(DEV)[1].item() == 2 and (out2 - out).abs().max() > 1e-3
    mine.eval()
    with torch.no_grad():
        e1, e2 = mine(fd, lens.to(DEV)), mine(fd, lens.to(DEV))
        assert torch.equal(e1, e2) and rel_err(e1.cpu(), ref(feat, lens)) < 5e-3
    assert mine._scb_dropout.get(DEV)[1].item() == 2


def test_cascaded_branch_train_mode_attention_dropout_vs_oracle_with_the_same_mask():
    from speechclip_b200 import ops
    from speechclip_b200.cascaded import SITE_MQ_ATTN
    from tests.test_cascaded_gpu import _branch_pair, _check_branch_grads
    p = 0.1
    cfg, model, oracle = _branch_pair("tiny", dropout=p)
    model.train()
    B, T, d, K = 12, 19, 64, 8
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(B, T, d, generator=g).half().float()
    lens = torch.randint(8, T + 1, (B,), generator=g)
    lens[0] = T
    cb, ob = model.cascaded_branch, oracle.cascaded_branch
    fd = feat.to(DEV).requires_grad_(True)
    mfeat, mvq, mkw = cb(fd, lens.to(DEV))
    state = cb._scb_dropout.get(DEV).clone()
    heads = cb.self_att.nhead
    mask = ops.dropout_mask(state, SITE_MQ_ATTN, p, B * heads * K * (T + K)).view(B, heads, K, T + K).cpu()
    w = torch.randn(mfeat.shape, generator=g)
    (mfeat * w.to(DEV)).sum().backward()
    # the oracle on its own first: the ids it selects are the CUDA path's, except where its two best scores are within the fp16
    # rounding of the K / V operands of each other; downstream of the selection the oracle is held to the CUDA path's ids
    from tests.test_cascaded_gpu import _assert_same_ids
    with torch.no_grad():
        collect = {}
        _, fvq, _ = ob(feat, lens, training=True, attn_mask_rows=mask, collect=collect)
    mine = mvq["targets"].cpu().view(-1)
    ref_scores = collect["cos"].detach().view(mine.numel(), -1).clone()
    ref_scores[:, [0, 2, 3]] = float("-inf")
    _assert_same_ids(mine, fvq["targets"].view(-1), ref_scores, tie=2e-3, min_agree=0.95)
    fr = feat.clone().requires_grad_(True)
    ofeat, ovq, okw = ob(fr, lens, training=True, attn_mask_rows=mask, force_idx=mine)
    (ofeat * w).sum().backward()
    assert rel_err(mfeat.detach().cpu(), ofeat.detach()) < 1e-2
    with torch.no_grad():
        plain, _, _ = ob(feat, lens, training=True)
    assert rel_err(plain, ofeat.detach()) > 2e-2   # the mask matters
    n = _check_branch_grads(dict(cb.named_parameters()), dict(ob.named_parameters()), 4e-2)
    assert n == 11
    assert rel_err(fd.grad.cpu(), fr.grad) < 4e-2
