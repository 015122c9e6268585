"""Every kernel behind the C ABI against a plain torch fp32 evaluation of the same op (and the oracle / golden fixtures
where the reference pins the op).  Tolerances are stated per test."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def gen(seed):
    return torch.Generator(device=DEV).manual_seed(seed)


def randn(*shape, seed=0, scale=1.0):
    return scale * torch.randn(*shape, device=DEV, generator=gen(seed))


# ------------------------------------------------------------------------------------------------ row kernels
@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
@pytest.mark.parametrize("d,act", [(512, 0), (768, 0), (1024, 1), (64, 2)])
def test_layernorm_fwd(dt, d, act):
    from speechclip_b200 import ops
    rows = 333
    x = randn(rows, d, seed=1).to(dt)
    g, b = 1 + 0.1 * randn(d, seed=2), 0.1 * randn(d, seed=3)
    y32 = torch.empty(rows, d, device=DEV)
    y16 = torch.empty(rows, d, device=DEV, dtype=torch.float16)
    stats = torch.empty(rows, 2, device=DEV)
    ops.layernorm(x, g, b, y32=y32, y16=y16, stats=stats, act=act)
    ref = F.layer_norm(x.float(), (d,), g, b, 1e-5)
    if act == 1:
        ref = F.gelu(ref)
    elif act == 2:
        ref = ref * torch.sigmoid(1.702 * ref)
    assert (y32 - ref).abs().max() < 2e-5
    assert (y16.float() - ref).abs().max() < 4e-3
    assert torch.allclose(stats[:, 0], x.float().mean(1), atol=1e-5)
    assert torch.allclose(stats[:, 1], (x.float().var(1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-4)


def test_layernorm_strided_rows_and_inplace_f16():
    from speechclip_b200 import ops
    B, L, d = 7, 5, 64
    x = randn(B, L, d, seed=4)
    g, b = 1 + 0.1 * randn(d, seed=5), 0.1 * randn(d, seed=6)
    y16 = torch.empty(B, d, device=DEV, dtype=torch.float16)
    ops.layernorm(x, g, b, y16=y16, rows=B, d=d, x_ld=L * d, y_ld=d)  # row 0 of every batch entry (ln_post on [CLS])
    assert (y16.float() - F.layer_norm(x[:, 0], (d,), g, b)).abs().max() < 4e-3
    h = randn(40, 512, seed=7).half()
    ref = F.gelu(F.layer_norm(h.float(), (512,), g.new_ones(512), g.new_zeros(512)))
    ops.layernorm(h, None, None, y16=h, rows=40, d=512, act=1)
    assert (h.float() - ref).abs().max() < 4e-3


def test_layernorm_bwd():
    from speechclip_b200 import ops
    rows, d = 257, 768
    x = randn(rows, d, seed=8).requires_grad_()
    g = (1 + 0.1 * randn(d, seed=9)).requires_grad_()
    b = (0.1 * randn(d, seed=10)).requires_grad_()
    dy = randn(rows, d, seed=11)
    F.layer_norm(x, (d,), g, b).backward(dy)
    y32, stats = torch.empty(rows, d, device=DEV), torch.empty(rows, 2, device=DEV)
    ops.layernorm(x.detach(), g.detach(), b.detach(), y32=y32, stats=stats)
    dx, dg, db = torch.empty(rows, d, device=DEV), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    ops.layernorm_bwd(dy, x.detach(), stats, g.detach(), dx, dg, db)
    assert (dx - x.grad).abs().max() < 2e-5
    assert (dg - g.grad).abs().max() < 2e-4 and (db - b.grad).abs().max() < 2e-4


def test_l2norm_fwd_bwd():
    from speechclip_b200 import ops
    x = randn(300, 512, seed=12).requires_grad_()
    dy = randn(300, 512, seed=13)
    ref = x / x.norm(dim=-1, keepdim=True)
    ref.backward(dy)
    y, n, dx = torch.empty(300, 512, device=DEV), torch.empty(300, device=DEV), torch.empty(300, 512, device=DEV)
    ops.l2norm(x.detach(), y, n)
    ops.l2norm_bwd(dy, y, n, dx)
    assert (y - ref).abs().max() < 1e-6 and (dx - x.grad).abs().max() < 1e-5


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("L,d", [(13, 768), (25, 1024), (3, 64)])
@pytest.mark.parametrize("hdt", [torch.float32, torch.float16])
def test_weighted_sum_fwd_bwd(normalize, L, d, hdt, golden):
    from oracle import speechclip as osc
    from speechclip_b200 import ops
    B, T = 3, 17
    h = randn(L, B * T, d, seed=14).to(hdt)  # fp16: the hidden states of the post-LN tower (the oracle sees the same rounded values)
    w = (0.5 * randn(L, seed=15)).requires_grad_()
    ref = osc.weighted_sum(w, list(h.float().view(L, B, T, d)), normalize)
    out32 = torch.empty(B * T, d, device=DEV)
    src16 = torch.zeros(B, T + 1, d, device=DEV, dtype=torch.float16)
    ops.weighted_sum(h, w.detach(), normalize, out32=out32, out16=src16, rows_per_batch=T, out16_batch_stride=(T + 1) * d, out16_row0=1)
    assert (out32.view(B, T, d) - ref).abs().max() < 1e-5
    assert (src16[:, 1:].float() - ref).abs().max() < 4e-3 and src16[:, 0].abs().max() == 0
    # backward through a strided dout (rows 1.. of a [B, T+1, d] buffer)
    dfull = randn(B, T + 1, d, seed=16)
    ref.backward(dfull[:, 1:])
    gw, scratch = torch.zeros(L, device=DEV), torch.empty(64, device=DEV)
    ops.weighted_sum_bwd(h, w.detach(), normalize, dfull[:, 1:], T, (T + 1) * d, 0, scratch, gw, 1.0)
    assert (gw - w.grad).abs().max() < 2e-4 * max(1.0, w.grad.abs().max().item())
    # reference-pinned fixture
    z = golden("ref_weighted_sum.npz")
    hid = torch.from_numpy(z["hidden"]).to(DEV)
    Lz, Bz, Tz, dz = hid.shape
    o = torch.empty(Bz * Tz, dz, device=DEV)
    ops.weighted_sum(hid.view(Lz, Bz * Tz, dz).contiguous(), torch.from_numpy(z["weights"]).to(DEV), normalize, out32=o)
    assert (o.view(Bz, Tz, dz).cpu() - torch.from_numpy(z[f"out_norm{int(normalize)}"])).abs().max() < 1e-5


def test_small_row_helpers():
    from speechclip_b200 import ops
    x, bias, res = randn(50, 96, seed=17), randn(96, seed=18), randn(50, 96, seed=19)
    pre, y = torch.empty(50, 96, device=DEV), torch.empty(50, 96, device=DEV)
    ops.rows_bias_act(x, bias, res, 96, 1, pre, y)
    assert torch.allclose(pre, x + bias + res, atol=1e-6) and torch.allclose(y, F.gelu(x + bias + res), atol=1e-6)
    ops.rows_bias_act(x, None, res[:1], 0, 0, None, y)
    assert torch.allclose(y, x + res[:1], atol=1e-6)
    p = randn(1000, seed=20).requires_grad_()
    dy = randn(1000, seed=21)
    F.gelu(p).backward(dy)
    dx = torch.empty(1000, device=DEV)
    ops.gelu_bwd(dy, p.detach(), dx)
    assert (dx - p.grad).abs().max() < 1e-6
    for dt in (torch.float32, torch.bfloat16):
        m = randn(1000, 70, seed=22).to(dt)
        out = torch.ones(70, device=DEV)
        ops.column_sum(m, out, beta=1.0)
        assert (out - (1 + m.float().sum(0))).abs().max() < 1e-3
        ops.column_sum(m, out)
        assert (out - m.float().sum(0)).abs().max() < 1e-3
    a = randn(37, 130, seed=23)
    t = torch.empty(130, 40, device=DEV, dtype=torch.bfloat16)
    ops.transpose(a, t[:, :37])
    assert (t[:, :37].float() - a.t()).abs().max() < 2e-2
    c = torch.empty(37, 130, device=DEV, dtype=torch.float16)
    ops.cast_rows(a, c)
    assert (c.float() - a).abs().max() < 2e-3
    o = torch.zeros(4, 3, 8, device=DEV)
    ops.broadcast_row(a[0, :8].contiguous(), a[1, :8].contiguous(), o, 24, 4, 8)
    assert torch.allclose(o[:, 0], (a[0, :8] + a[1, :8]).expand(4, 8)) and o[:, 1:].abs().max() == 0
    lens = torch.tensor([0, 5, 319, 400], device=DEV)
    kv = torch.empty(4, device=DEV, dtype=torch.int32)
    ops.lengths_to_i32(lens, 1, 320, kv)
    assert kv.tolist() == [1, 6, 320, 320]


def test_sgemm_strided():
    from speechclip_b200 import ops
    a, b = randn(70, 45, seed=24), randn(33, 45, seed=25)
    c = torch.empty(70, 33, device=DEV)
    ops.sgemm(a, b, c)
    assert (c - a @ b.t()).abs().max() < 1e-4
    at, bt = randn(45, 70, seed=26), randn(45, 33, seed=27)
    c0 = randn(70, 33, seed=28)
    c = c0.clone()
    ops.sgemm(at.t(), bt.t(), c, alpha=0.5, beta=1.0)
    assert (c - (c0 + 0.5 * at.t() @ bt)).abs().max() < 1e-4
    v, u = randn(64, seed=29), randn(48, seed=30)
    o = torch.empty(64, 48, device=DEV)
    ops.sgemm(v.view(64, 1), u.view(48, 1), o)
    assert (o - torch.outer(v, u)).abs().max() < 1e-6


@pytest.mark.parametrize("M", [1, 3, 8])
def test_sgemm_skinny_rows(M):
    """M <= 8 takes the dot-product kernels: B contiguous in k (q = cls Wq^T) or in n (dcls += dq Wq), alpha / beta honoured."""
    from speechclip_b200 import ops
    K, N = 768, 200
    a, w = randn(M, K, seed=31), randn(N, K, seed=32, scale=K ** -0.5)
    c = torch.empty(M, N, device=DEV)
    ops.sgemm(a, w, c)
    assert (c - a @ w.t()).abs().max() < 1e-4
    wt = randn(K, N, seed=33, scale=K ** -0.5)  # b = wt.t(): element (n, k) at k * N + n
    c0 = randn(M, N, seed=34)
    c = c0.clone()
    ops.sgemm(a, wt.t(), c, alpha=2.0, beta=1.0)
    assert (c - (c0 + 2.0 * a @ wt)).abs().max() < 1e-4


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("hd,heads,T", [(64, 12, 319), (64, 12, 50), (16, 4, 12), (96, 8, 320), (128, 8, 320), (64, 8, 77), (32, 2, 130)])
@pytest.mark.parametrize("mode", ["full", "keypad", "causal"])
def test_attention_fwd(hd, heads, T, mode):
    from speechclip_b200 import ops
    B, d = 3, heads * hd
    qkv = randn(B, T, 3 * d, seed=31, scale=0.7).half()
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    out = torch.empty(B, T, d, device=DEV, dtype=torch.float16)
    kv_len = None
    lens = [T, max(1, T // 3), max(1, T - 1)]
    if mode == "keypad":
        kv_len = torch.tensor(lens, device=DEV, dtype=torch.int32)
    ops.attention(q, k, v, out, heads, hd ** -0.5, kv_len, causal=(mode == "causal"))
    qf, kf, vf = (t.float().view(B, T, heads, hd).transpose(1, 2) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * hd ** -0.5
    if mode == "keypad":
        pad = torch.arange(T, device=DEV)[None] >= torch.tensor(lens, device=DEV)[:, None]
        s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    if mode == "causal":
        s = s + torch.full((T, T), float("-inf"), device=DEV).triu_(1)
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, T, d)
    err = (out.float() - ref).abs().max().item()
    assert err < 4e-3, err  # fp16 P and fp16 output


def test_attention_tcgen05_kernel_all_supported_shapes(monkeypatch):
    """The tcgen05/TMEM attention on every shape class it supports, including the short ones the dispatcher normally leaves to
    the mma.sync kernel (run in a subprocess: the dispatch knob is read once per process)."""
    import os
    import subprocess
    import sys
    code = r"""
import torch
from speechclip_b200 import ops
for (B, T, heads, causal, lens) in [(3, 319, 12, False, [319, 100, 318]), (2, 50, 12, False, None), (2, 77, 8, True, None), (2, 257, 16, False, None), (5, 320, 4, False, [320, 1, 64, 65, 200])]:
    d = heads * 64
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = (0.7 * torch.randn(B, T, 3 * d, device="cuda", generator=g)).half()
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    out = torch.empty(B, T, d, device="cuda", dtype=torch.float16)
    kv_len = torch.tensor(lens, device="cuda", dtype=torch.int32) if lens else None
    ops.attention(q, k, v, out, heads, 0.125, kv_len, causal=causal)
    qf, kf, vf = (t.float().view(B, T, heads, 64).transpose(1, 2) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * 0.125
    if lens:
        pad = torch.arange(T, device="cuda")[None] >= torch.tensor(lens, device="cuda")[:, None]
        s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    if causal:
        s = s + torch.full((T, T), float("-inf"), device="cuda").triu_(1)
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, T, d)
    err = (out.float() - ref).abs().max().item()
    assert err < 4e-3, (T, heads, causal, err)
print("ok")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SCB_ATTN_TC="2")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("hd,heads,Tk", [(96, 8, 320), (16, 4, 13), (128, 8, 320), (8, 8, 12)])
def test_cls_attention_fwd_bwd(hd, heads, Tk):
    from speechclip_b200 import ops
    B, d = 5, heads * hd
    kv = randn(B, Tk, 2 * d, seed=32, scale=0.8).half()
    q = randn(1, d, seed=33).requires_grad_()
    lens = torch.tensor([Tk, 1, Tk // 2, Tk - 1, 3][:B], device=DEV, dtype=torch.int32)
    probs, ctx = torch.empty(B, heads, Tk, device=DEV), torch.empty(B, d, device=DEV)
    ops.cls_attention_fwd(q.detach(), kv, 0, d, lens, heads, hd, hd ** -0.5, probs, ctx)
    kvf = kv.float().requires_grad_()
    kf = kvf[..., :d].view(B, Tk, heads, hd).transpose(1, 2)
    vf = kvf[..., d:].view(B, Tk, heads, hd).transpose(1, 2)
    s = (q.view(1, heads, 1, hd) * hd ** -0.5) @ kf.transpose(-1, -2)
    pad = torch.arange(Tk, device=DEV)[None] >= lens[:, None]
    s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    ref = (p @ vf).transpose(1, 2).reshape(B, d)
    assert (ctx - ref).abs().max() < 1e-4 and (probs - p.squeeze(2)).abs().max() < 1e-5
    dctx = randn(B, d, seed=34)
    ref.backward(dctx)
    dkv = torch.empty(B, Tk, 2 * d, device=DEV, dtype=torch.bfloat16)
    dq = torch.zeros(d, device=DEV)
    ops.cls_attention_bwd(q.detach(), kv, 0, d, lens, heads, hd, hd ** -0.5, probs, dctx, dkv, dq)
    assert (dq - q.grad.view(-1)).abs().max() < 2e-4 * max(1.0, q.grad.abs().max().item())
    scale = kvf.grad.abs().max().item()
    assert (dkv.float() - kvf.grad).abs().max() < 1e-2 * scale  # bf16 storage of the gradient


# ------------------------------------------------------------------------------------------------ front end
def test_frame_lengths_matches_oracle_rules():
    from oracle import hubert as oh
    from speechclip_b200 import ops
    lens = [102400, 48000, 48160, 400, 35000, 160 * 3, 101999, 800]
    tw = 102400
    T = oh.conv_out_length(tw)
    wl = torch.tensor(lens, device=DEV)
    B = len(lens)
    ints = torch.empty(4, B, device=DEV, dtype=torch.int32)
    fl64 = torch.empty(B, device=DEV, dtype=torch.int64)
    ops.frame_lengths(wl, B, tw, 0, T, 320, None, ints[0], ints[1], ints[2], ints[3], fl64)
    pad = ~(torch.arange(tw)[None] < torch.tensor(lens)[:, None])
    valid = (~oh.HubertModel.frame_padding_mask(T, pad)).sum(1)
    assert ints[2].cpu().tolist() == valid.tolist()
    assert fl64.cpu().tolist() == oh.feat_lengths(lens, T).tolist() == ints[3].cpu().tolist()
    assert ints[1].cpu().tolist() == lens and ints[0].abs().max() == 0
    # training crop: len > max_audio_len -> window of max_audio_len at floor(u * (len - max + 1))
    wl2 = torch.tensor([200000, 5000, 102401], device=DEV)
    u = torch.tensor([0.5, 0.9, 0.999], device=DEV)
    ops.frame_lengths(wl2, 3, 102400, 102400, T, 320, u, ints[0, :3], ints[1, :3], ints[2, :3], ints[3, :3], fl64[:3])
    assert ints[1, :3].cpu().tolist() == [102400, 5000, 102400]
    assert ints[0, :3].cpu().tolist() == [int(0.5 * (200000 - 102400 + 1)), 0, 1]


def test_wav_prepare_crop_and_normalize():
    from speechclip_b200 import ops
    B, Tmax, tw = 3, 5000, 3000
    wav = randn(B, Tmax, seed=35, scale=0.1) + 0.02
    off = torch.tensor([100, 0, 2000], device=DEV, dtype=torch.int32)
    ln = torch.tensor([3000, 1234, 3000], device=DEV, dtype=torch.int32)
    out = torch.empty(B, tw, device=DEV)
    ops.wav_prepare(wav, off, ln, tw, False, None, out)
    for b in range(B):
        o, l = int(off[b]), int(ln[b])
        assert torch.equal(out[b, :l], wav[b, o:o + l]) and out[b, l:].abs().max().item() == 0 if l < tw else True
    stats = torch.empty(2 * B, device=DEV)
    ops.wav_prepare(wav, off, ln, tw, True, stats, out)
    for b in range(B):
        o, l = int(off[b]), int(ln[b])
        seg = wav[b, o:o + l]
        assert (out[b, :l] - F.layer_norm(seg, seg.shape)).abs().max() < 2e-5


@pytest.mark.parametrize("n_samples", [4000, 16000, 10 + 5 * 63])
def test_conv0_groupnorm_gelu(n_samples):
    from speechclip_b200 import ops
    B = 3
    wav = randn(B, n_samples, seed=36, scale=0.1)
    wav[1, n_samples // 2:] = 0  # zero padding participates in the GroupNorm statistics, like fairseq
    w = randn(512, 10, seed=37, scale=math.sqrt(2 / 10))
    gamma, beta = 1 + 0.1 * randn(512, seed=38), 0.1 * randn(512, seed=39)
    T = (n_samples - 10) // 5 + 1
    out = torch.empty(B, T, 512, device=DEV, dtype=torch.float16)
    scratch = torch.empty(ops.conv0_scratch_bytes(B), device=DEV, dtype=torch.uint8)
    ops.conv0_groupnorm_gelu(wav, n_samples, w, None, gamma, beta, 1e-5, out, T * 512, scratch)
    ref = F.gelu(F.group_norm(F.conv1d(wav[:, None], w[:, None], stride=5), 512, gamma, beta, 1e-5)).transpose(1, 2)
    err = (out.float() - ref).abs().max().item()
    assert err < 6e-3, err


@pytest.mark.parametrize("with_bias", [False, True])
@pytest.mark.parametrize("n", [4000, 10 + 5 * 130])
def test_conv0_layernorm_gelu(with_bias, n):
    from speechclip_b200 import ops
    B = 2
    wav = randn(B, n, seed=40)
    bias = 0.3 * randn(512, seed=44) if with_bias else None
    w = randn(512, 10, seed=41, scale=math.sqrt(2 / 10))
    gamma, beta = 1 + 0.1 * randn(512, seed=42), 0.1 * randn(512, seed=43)
    T = (n - 10) // 5 + 1
    out = torch.empty(B, T, 512, device=DEV, dtype=torch.float16)
    scratch = torch.empty(ops.conv0_scratch_bytes(B), device=DEV, dtype=torch.uint8)
    ops.conv0_layernorm_gelu(wav, n, w, bias, gamma, beta, 1e-5, out, T * 512, scratch)
    ref = F.gelu(F.layer_norm(F.conv1d(wav[:, None], w[:, None], bias, stride=5).transpose(1, 2), (512,), gamma, beta, 1e-5))
    err = (out.float() - ref).abs().max().item()
    assert err < 6e-3, err


@pytest.mark.parametrize("k,T_in", [(3, 799), (3, 400), (2, 159), (2, 80)])
def test_conv_tap_walk_gemm_matches_conv1d(k, T_in):
    """conv1..6: stride-2 Conv1d over channel-last rows as a GEMM whose A rows are pairs of frames (no im2col)."""
    from speechclip_b200 import ops
    B, C = 3, 512
    x = randn(B, T_in, C, seed=44, scale=0.5).half()
    buf = torch.zeros(B * T_in * C + 2048, device=DEV, dtype=torch.float16)
    buf[:B * T_in * C] = x.reshape(-1)
    w = randn(C, C, k, seed=45, scale=math.sqrt(2 / (C * k))).half()
    wk = w.permute(0, 2, 1).reshape(C, k * C).contiguous()
    T_out = (T_in - k) // 2 + 1
    out = torch.empty(B, T_out, C, device=DEV, dtype=torch.float16)
    ops.gemm_raw(a=buf, a_inner=1024, a_rows=(T_in + 1) // 2, a_row_stride=1024, a_batch_stride=T_in * C, batch=B, m_per_batch=T_out,
                 w=wk, n=C, k=k * C, kb_per_tap=16, tap_row_shift=1, out=out, ldc=C, out_batch_stride=T_out * C, act=1)
    ref = F.gelu(F.conv1d(x.float().transpose(1, 2), w.float(), stride=2)).transpose(1, 2)
    err = (out.float() - ref).abs().max().item()
    assert err < 6e-3, err


@pytest.mark.parametrize("d,G,K,T", [(768, 16, 128, 319), (64, 4, 16, 12), (1024, 16, 128, 49)])
def test_positional_conv_grouped_gemm(d, G, K, T):
    from speechclip_b200 import ops
    B, cpg = 2, d // G
    x = randn(B, T, d, seed=46, scale=0.5)
    valid = torch.tensor([T, max(1, T // 2)], device=DEV, dtype=torch.int32)
    w = randn(d, cpg, K, seed=47, scale=math.sqrt(1.0 / (cpg * K))).half().float()
    bias = 0.1 * randn(d, seed=48)
    wp = torch.zeros(G, cpg, K, 64, device=DEV)
    wp[..., :cpg] = w.view(G, cpg, cpg, K).permute(0, 1, 3, 2)
    wp = wp.reshape(G, cpg, K * 64).half().contiguous()
    rows_pad = T + K
    xpad = torch.full((B, rows_pad, G * 64), float("nan"), device=DEV, dtype=torch.float16)  # the pack kernel writes every element
    xm = x.clone()
    ops.posconv_pack(xm, valid, xpad, B, T, d, G, K // 2, rows_pad)
    mask = torch.arange(T, device=DEV)[None] >= valid[:, None]
    x0 = x.masked_fill(mask[:, :, None], 0.0)
    assert torch.equal(xm, x0)
    out = torch.empty(B, T, d, device=DEV)
    ops.gemm_raw(a=xpad, a_inner=G * 64, a_rows=rows_pad, a_row_stride=G * 64, a_batch_stride=rows_pad * G * 64, batch=B, m_per_batch=T,
                 w=wp, n=cpg, k=K * 64, groups=G, b_group_stride=cpg * K * 64, kb_per_tap=1, tap_row_shift=1, a_group_cols=64, out=out,
                 ldc=d, out_batch_stride=T * d, out_group_cols=cpg, bias=bias, act=1, residual=xm)
    pc = F.conv1d(x0.half().float().transpose(1, 2), w, bias, padding=K // 2, groups=G)[:, :, :-1]
    ref = x0 + F.gelu(pc).transpose(1, 2)
    err = (out - ref).abs().max().item()
    assert err < 3e-3, err


@pytest.mark.parametrize("S,P,W", [(224, 32, 768), (32, 16, 64), (224, 14, 1024)])
def test_patchify_gemm_matches_conv2d(S, P, W):
    from speechclip_b200 import ops
    B, G = 3, S // P
    img = randn(B, 3, S, S, seed=49)
    w = randn(W, 3, P, P, seed=50, scale=(3 * P * P) ** -0.5).half().float()
    kk = 3 * P * P
    ldk = (kk + 7) // 8 * 8
    wp = torch.zeros(W, ldk, device=DEV)
    wp[:, :kk] = w.view(W, kk)
    wp = wp.half()
    pos, cls = 0.1 * randn(G * G + 1, W, seed=51), randn(W, seed=52)
    patches = torch.empty(B * G * G, ldk, device=DEV, dtype=torch.float16)
    ops.patchify(img, patches, P, ldk)
    L = G * G + 1
    tok = torch.empty(B, L, W, device=DEV)
    ops.gemm_raw(a=patches, a_inner=ldk, a_rows=G * G, a_row_stride=ldk, a_batch_stride=G * G * ldk, batch=B, m_per_batch=G * G, w=wp,
                 n=W, k=ldk, out=tok, out_offset=W, ldc=W, out_batch_stride=L * W, residual=pos, residual_offset=W, residual_ld=W,
                 residual_batch_stride=0)
    ops.broadcast_row(cls, pos, tok, L * W, B, W)
    x = F.conv2d(img.half().float(), w, stride=P).reshape(B, W, -1).permute(0, 2, 1)
    ref = torch.cat([cls.expand(B, 1, W), x], 1) + pos
    err = (tok - ref).abs().max().item()
    assert err < 3e-3, err


# ------------------------------------------------------------------------------------------------ loss / optimiser / retrieval
def _loss_call(a, b, ids, log_mult=None, mult=1 / 0.07, margin=0.0, dcl=False, a2b=True, b2a=True, upstream=None):
    from speechclip_b200 import ops
    B, D = a.shape
    loss = torch.empty((), device=DEV)
    scratch = torch.empty(ops.infonce_scratch_bytes(B), device=DEV, dtype=torch.uint8)
    dA, dB = torch.empty_like(a), torch.empty_like(b)
    dT = torch.zeros((), device=DEV) if log_mult is not None else None
    logits = torch.empty(B, B, device=DEV)
    ops.infonce(a, b, ids, log_mult, mult, margin, dcl, a2b, b2a, scratch, phase=1, loss=loss, logits_out=logits)
    ops.infonce(a, b, ids, log_mult, mult, margin, dcl, a2b, b2a, scratch, phase=2, upstream_dev=upstream, dA=dA, dB=dB, dlog_mult=dT)
    return loss, dA, dB, dT, logits


@pytest.mark.parametrize("tag", ["small", "mid"])
def test_infonce_matches_reference_fixture(golden, tag):
    z = golden(f"ref_loss_{tag}.npz")
    T = lambda k: torch.from_numpy(z[k]).to(DEV)
    a, b, ids = T("a"), T("b"), T("ids")
    loss, dA, dB, _, _ = _loss_call(a, b, ids)
    assert abs(loss.item() - float(z["loss"])) < 2e-5
    assert (dA - T("da")).abs().max() < 2e-6 and (dB - T("db")).abs().max() < 2e-6
    tp = T("temp_param")
    loss_t, dA_t, _, dT, _ = _loss_call(a, b, ids, log_mult=tp)
    assert abs(loss_t.item() - float(z["loss_t"])) < 2e-5
    assert abs(dT.item() - float(z["dtemp"])) < 1e-4 and (dA_t - T("da_t")).abs().max() < 2e-6
    assert abs(_loss_call(a, b, None)[0].item() - float(z["loss_noid"])) < 2e-5


@pytest.mark.parametrize("B,D", [(256, 512), (300, 768), (1000, 64)])
@pytest.mark.parametrize("variant", ["plain", "margin_dcl", "a2b_only"])
def test_infonce_vs_oracle(B, D, variant):
    from oracle import speechclip as osc
    a = F.normalize(randn(B, D, seed=53), dim=-1)
    b = F.normalize(randn(B, D, seed=54), dim=-1)
    ids = (torch.randperm(B, device=DEV, generator=gen(55)) // 5).contiguous()
    kw = dict(plain={}, margin_dcl=dict(margin=0.2, dcl=True), a2b_only=dict(b2a=False))[variant]
    a_c, b_c = a.detach().cpu().requires_grad_(), b.detach().cpu().requires_grad_()
    ref, ref_logits = osc.masked_contrastive_loss(a_c, b_c, ids.cpu(), 1 / 0.07, return_logits=True, **kw)
    (ref * 0.5).backward()
    up = torch.tensor(0.5, device=DEV)
    loss, dA, dB, _, logits = _loss_call(a.detach(), b.detach(), ids, upstream=up, **kw)
    assert abs(loss.item() - ref.item()) < 3e-5 * max(1.0, abs(ref.item()))
    assert (logits.cpu() - ref_logits.detach()).abs().max() < 2e-5
    gmax = a_c.grad.abs().max().item()
    assert (dA.cpu() - a_c.grad).abs().max() < 2e-4 * gmax and (dB.cpu() - b_c.grad).abs().max() < 2e-4 * gmax
    # bit-exact top-1 retrieval indices on the logits (north_star)
    assert torch.equal(logits.argmax(1).cpu(), ref_logits.argmax(1)) and torch.equal(logits.argmax(0).cpu(), ref_logits.argmax(0))


def test_infonce_global_batch_4096_tensor_path():
    """BASELINE config 5: 4096 x 768 global contrastive — the three 25.8 GFLOP contractions run as TF32 on the tensor cores.
    Tolerance: logits within 1e-3 of the logit scale (north_star), loss 1e-3 relative, gradients 2 % of their max."""
    from oracle import speechclip as osc
    B, D = 4096, 768
    a = F.normalize(randn(B, D, seed=61), dim=-1)
    b = F.normalize(randn(B, D, seed=62) + 0.5 * a, dim=-1)  # correlated pairs: a realistic, peaked similarity matrix
    ids = (torch.randperm(B, device=DEV, generator=gen(63)) // 5).contiguous()
    a_c, b_c = a.cpu().requires_grad_(), b.cpu().requires_grad_()
    ref, ref_logits = osc.masked_contrastive_loss(a_c, b_c, ids.cpu(), 1 / 0.07, return_logits=True)
    ref.backward()
    loss, dA, dB, _, logits = _loss_call(a, b, ids)
    assert (logits.cpu() - ref_logits.detach()).abs().max() < 1e-3 / 0.07
    assert abs(loss.item() - ref.item()) < 1e-3 * abs(ref.item())
    gmax = a_c.grad.abs().max().item()
    assert (dA.cpu() - a_c.grad).abs().max() < 2e-2 * gmax and (dB.cpu() - b_c.grad).abs().max() < 2e-2 * gmax


def test_adam_step_matches_torch_adam_with_clipping():
    from speechclip_b200 import ops
    n = 100003
    p0, g = randn(n, seed=56), randn(n, seed=57, scale=0.3)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-2)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    sumsq = torch.zeros(1, device=DEV, dtype=torch.float64)
    for step in range(1, 4):
        gi = g * step
        ref.grad = gi.clone()
        torch.nn.utils.clip_grad_norm_([ref], 4.0)
        opt.step()
        ops.adam_step(p, gi, m, v, sumsq, 1.0, 4.0, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step)
        assert (p - ref.data).abs().max() < 2e-6, step


def test_retrieval_rank_matches_oracle_and_fixture(golden):
    from avssl.module import mutualRetrieval
    from oracle import speechclip as osc
    z = golden("ref_retrieval.npz")
    s = torch.from_numpy(z["score"]).to(DEV)
    ab, ba = torch.from_numpy(z["ab"]).to(DEV), torch.from_numpy(z["ba"]).to(DEV)
    rAB, rBA, rM = mutualRetrieval(s, s.t().contiguous(), ab, ba, [1, 5, 10])
    for i, k in enumerate((1, 5, 10)):
        assert abs(rAB[f"recall@{k}"] - z["rAB"][i]) < 1e-4 and abs(rBA[f"recall@{k}"] - z["rBA"][i]) < 1e-4
        assert abs(rM[f"recall@{k}"] - z["rM"][i]) < 1e-4
    nA, nB = 5000, 1000
    sc = randn(nA, nB, seed=58)
    a_ans = torch.arange(nA, device=DEV) // 5
    b_ans = torch.arange(nB, device=DEV)
    mine = mutualRetrieval(sc, sc.t().contiguous(), a_ans, b_ans, [1, 5, 10])
    ref = osc.mutual_retrieval(sc.cpu(), sc.t().contiguous().cpu(), a_ans.cpu(), b_ans.cpu(), [1, 5, 10])
    for m_, r_ in zip(mine, ref):
        for k in r_:
            assert abs(m_[k] - r_[k]) < 1e-4
    from speechclip_b200 import ops
    top1 = torch.empty(nA, device=DEV, dtype=torch.int32)
    ops.retrieval_rank(sc, None, None, None, top1)
    assert torch.equal(top1.long(), sc.argmax(1))


# ------------------------------------------------------------------------------------------------ transposes
@pytest.mark.parametrize("rows,cols,out_pad", [(64, 64, 0), (4608, 1536, 256), (328, 72, 8), (81920 // 8, 768, 0), (100, 36, 0), (7, 5, 0)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_transpose_16bit(rows, cols, out_pad, dt):
    """dst[c, r] = src[r, c]: the vectorised 64 x 64 kernel (rows, cols, leading dimensions multiples of 8) and the generic one;
    the output may be a column slice of a wider buffer (the wgrad operand buffers are padded to the split-K length)."""
    from speechclip_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(rows + cols)
    src_full = torch.randn(rows, cols + 8, device=DEV, generator=g).to(dt)
    src = src_full[:, :cols]                      # row stride cols + 8
    buf = torch.full((cols, rows + out_pad), 3.0, device=DEV, dtype=dt)
    ops.transpose(src, buf[:, :rows])
    torch.cuda.synchronize()
    assert torch.equal(buf[:, :rows], src.t())
    if out_pad:
        assert (buf[:, rows:] == 3.0).all()


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("d,act,affine", [(512, 1, True), (512, 0, False), (264, 2, True), (64, 1, True)])
def test_layernorm_16bit_rows_kernel(dt, d, act, affine):
    """16-bit in, 16-bit out, many rows (the HuBERT-large conv blocks run LayerNorm + GELU in place over B x T x 512): the
    grid-stride kernel with gamma / beta in registers and 16-byte accesses; in place and into a strided output."""
    from speechclip_b200 import ops
    rows = 5003
    x = randn(rows, d, seed=d + act).to(dt)
    g = 1 + 0.1 * randn(d, seed=2) if affine else None
    b = 0.1 * randn(d, seed=3) if affine else None
    ref = F.layer_norm(x.float(), (d,), g, b, 1e-5)
    if act == 1:
        ref = F.gelu(ref)
    elif act == 2:
        ref = ref * torch.sigmoid(1.702 * ref)
    tol = 6e-3 if dt == torch.float16 else 4e-2
    wide = torch.full((rows, d + 16), 5.0, device=DEV, dtype=dt)
    ops.layernorm(x, g, b, y16=wide[:, :d], rows=rows, d=d, x_ld=d, y_ld=d + 16, act=act)
    assert (wide[:, :d].float() - ref).abs().max() < tol
    assert (wide[:, d:] == 5.0).all()
    ops.layernorm(x, g, b, y16=x, rows=rows, d=d, act=act)   # in place
    assert (x.float() - ref).abs().max() < tol
