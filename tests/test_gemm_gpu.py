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
