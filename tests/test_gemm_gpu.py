"""tcgen05 GEMM vs a plain torch fp32 reference of the same contraction (on the same 16-bit-rounded operands)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias, act, residual, alpha=1.0):
    x = alpha * (a.float() @ w.float().t())
    if bias is not None:
        x = x + bias
    if act == 1:
        x = torch.nn.functional.gelu(x)
    elif act == 2:
        x = x * torch.sigmoid(1.702 * x)
    if residual is not None:
        x = x + residual.float()
    return x


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (300, 768, 768), (638, 2304, 768), (1000, 512, 3072),
                                   (77, 48, 128), (256, 128, 192), (4097, 3072, 768), (50, 8, 64)])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_gemm_plain(M, N, K, dt):
    from speechclip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(dt)
    out = ops.gemm(a, w, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = _ref(a, w, None, 0, None)
    err = (out - ref).abs().max().item()
    assert err < 2e-3, err


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("out_dtype,res_dtype", [(torch.float32, torch.float32), (torch.float16, torch.float16), (torch.float16, None)])
def test_gemm_epilogue(act, out_dtype, res_dtype):
    from speechclip_b200 import ops
    M, N, K = 700, 776, 320
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(res_dtype) if res_dtype else None
    out2 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    out = ops.gemm(a, w, bias=bias, act=act, residual=res, out_dtype=out_dtype, out2=out2, alpha=0.5)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias, act, res, 0.5)
    tol = 2e-3 if out_dtype == torch.float32 else 8e-3
    assert (out.float() - ref).abs().max().item() < tol
    assert (out2.float() - ref).abs().max().item() < 8e-3


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("M,N,K,bias", [(700, 776, 320, True), (20000, 2304, 768, True), (4097, 3072, 768, False),
                                        (9000, 264, 128, True)])
def test_gemm_tma_store_epilogue(act, M, N, K, bias):
    """16-bit output, no residual, no second output, N > 128: the epilogue leaves through TMA stores (32 x 64 boxes clipped by
    the tensor map at the M and N edges; single-CTA kernel below 74 tile pairs, CTA pairs above)."""
    from speechclip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K + act)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    guard = torch.full((M + 64, N), 7.0, device="cuda", dtype=torch.float16)
    out = guard[:M]
    ops.gemm(a, w, bias=b, act=act, out=out, alpha=1.5)
    torch.cuda.synchronize()
    ref = _ref(a, w, b, act, None, 1.5)
    err = (out.float() - ref).abs().max().item()
    assert err < 8e-3, err
    assert (guard[M:] == 7.0).all()  # rows past M are clipped, not written


@pytest.mark.parametrize("tap", [0, 1, 2])
def test_gemm_tma_store_batched_ragged_rows(tap):
    """Batched output whose rows per batch are not a multiple of the 32-row store box (the conv stack: [B][T_out][512]): a box
    that straddles the end of one batch must not spill into the next one."""
    from speechclip_b200 import ops
    B, T, C, N = (3, 4301, 256, 512) if tap == 2 else (5, 333, 256, 512)  # tap 2: >= 2048 rows per batch -> the tap walk runs on CTA pairs
    g = torch.Generator(device="cuda").manual_seed(11 + tap)
    x = torch.randn(B, T, C, device="cuda", generator=g).half()
    if tap:  # stride-2, k = 2 conv over channel-last rows: pairs of frames are one 2C-wide row
        T_out = T // 2
        w = (torch.randn(N, 2 * C, device="cuda", generator=g) / (2 * C) ** 0.5).half()
        out = torch.zeros(B, T_out, N, device="cuda", dtype=torch.float16)
        ops.gemm_raw(a=x, a_inner=2 * C, a_rows=(T + 1) // 2, a_row_stride=2 * C, a_batch_stride=T * C, batch=B, m_per_batch=T_out,
                     w=w, n=N, k=2 * C, kb_per_tap=2 * C // 64, tap_row_shift=1, out=out, ldc=N, out_batch_stride=T_out * N,
                     act=ops.ACT_GELU)
        ref = torch.nn.functional.gelu(x[:, :2 * T_out].reshape(B, T_out, 2 * C).float() @ w.float().t())
    else:
        w = (torch.randn(N, C, device="cuda", generator=g) / C ** 0.5).half()
        out = torch.zeros(B, T, N, device="cuda", dtype=torch.float16)
        ops.gemm_raw(a=x, a_inner=C, a_rows=T, a_row_stride=C, a_batch_stride=T * C, batch=B, m_per_batch=T, w=w, n=N, k=C,
                     out=out, ldc=N, out_batch_stride=T * N, act=ops.ACT_GELU)
        ref = torch.nn.functional.gelu(x.float() @ w.float().t())
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err < 8e-3, err


def test_gelu_h16_fit_error():
    """The sigmoid-polynomial erf-GELU of the 16-bit epilogues stays within 3e-5 of erf-GELU before the fp16 rounding: checked
    through a K = 64 identity contraction so the epilogue sees exactly the probe values."""
    from speechclip_b200 import ops
    n = 256
    xs = torch.linspace(-12.0, 12.0, 1 << 20, device="cuda").half()  # 16384 rows: enough tiles to stay on the 256-wide kernel
    # out[m, j] = a[m, :] . w[j, :] = probe value (m, j % 64): w = one-hot rows, a carries 64 probes per row
    rows = xs.numel() // 64
    a = xs.view(rows, 64).contiguous()
    w = torch.zeros(n, 64, device="cuda", dtype=torch.float16)
    w[torch.arange(n), torch.arange(n) % 64] = 1.0
    out = ops.gemm(a, w, act=ops.ACT_GELU)
    torch.cuda.synchronize()
    probe = a.float()[:, torch.arange(n, device="cuda") % 64]
    ref = torch.nn.functional.gelu(probe.double()).float()
    err = (out.float() - ref).abs()
    ulp = ref.abs().clamp_min(2.0 ** -14) * 2.0 ** -11  # half an fp16 ulp, roughly
    assert (err <= ulp + 3e-5).all(), (err - ulp).max().item()


@pytest.mark.parametrize("res_dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("M,N,K,bias", [(700, 776, 320, True), (20000, 768, 768, True), (4097, 3072, 256, False), (12800, 768, 3072, True)])
def test_gemm_tma_store_residual_epilogue(res_dtype, M, N, K, bias):
    """fp32 output = alpha * acc + bias + residual (out-proj / fc2), residual fp32 (pre-LN towers) or fp16 (post-LN hidden
    states): row-per-lane residual reads + TMA stores of 32 x 32 fp32 boxes, on single CTAs and on CTA pairs; also in place."""
    from speechclip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    res = torch.randn(M, N, device="cuda", generator=g).to(res_dtype)
    guard = torch.full((M + 64, N), 7.0, device="cuda", dtype=torch.float32)
    out = guard[:M]
    ops.gemm(a, w, bias=b, residual=res, out=out, alpha=0.5)
    torch.cuda.synchronize()
    ref = _ref(a, w, b, 0, res, 0.5)
    err = (out - ref).abs().max().item()
    assert err < 2e-3, err
    assert (guard[M:] == 7.0).all()
    if res_dtype == torch.float32:  # in place: out aliases the residual
        buf = res.clone()
        ops.gemm(a, w, bias=b, residual=buf, out=buf, alpha=0.5)
        torch.cuda.synchronize()
        assert (buf - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("M,N,K,act,out_dtype,res_dtype", [
    (10208, 768, 3072, 0, torch.float32, torch.float16),   # HuBERT fc2 at 32 utterances: 120 tile pairs on 74 SM pairs
    (10208, 768, 768, 0, torch.float32, torch.float16),    # out-proj
    (10208, 3072, 768, 1, torch.float16, None),            # fc1 (6.49 waves), GELU + TMA store
    (1600, 768, 3072, 0, torch.float32, torch.float32),    # CLIP ViT-B/32 fc2 at 32 images: 21 tile pairs -> every tile cut in ~3.5
    (1600, 3072, 768, 2, torch.float16, None),
    (12800, 768, 3072, 0, torch.float32, torch.float32),   # ViT fc2 at 256 images: 150 tile pairs = 2.03 waves
    (1000, 520, 4096, 0, torch.float32, None),             # ragged M / N edges inside split tiles, generic epilogue
])
def test_gemm_stream_k_tail(M, N, K, act, out_dtype, res_dtype):
    """With a workspace the CTA-pair kernel cuts the tiles of its last, partial wave along K over all SM pairs; partial sums
    meet in the owner's TMEM accumulator.  Same result as without (up to fp32 summation order), repeated launches through
    the same workspace stay correct (the arrival counters are reset by the owners), and the workspace tail is not overrun."""
    from speechclip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    b = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(res_dtype) if res_dtype else None
    nb = ops.gemm_workspace_bytes()
    scratch = torch.zeros(nb + 4096, device="cuda", dtype=torch.uint8)
    scratch[nb:] = 0x5A
    ref = _ref(a, w, b, act, res, 0.5)
    plain = ops.gemm(a, w, bias=b, act=act, residual=res, out_dtype=out_dtype, alpha=0.5)
    tol = 2e-3 if out_dtype == torch.float32 else 8e-3
    for _ in range(3):
        out = torch.full((M, N), 9.0, device="cuda", dtype=out_dtype)
        ops.gemm(a, w, bias=b, act=act, residual=res, out=out, alpha=0.5, scratch=scratch[:nb])
        torch.cuda.synchronize()
        assert (out.float() - ref).abs().max().item() < tol
        assert (out.float() - plain.float()).abs().max().item() < (1e-4 if out_dtype == torch.float32 else 4e-3)
        assert (scratch[:4096] == 0).all()      # counters back to zero
    assert (scratch[nb:] == 0x5A).all()
