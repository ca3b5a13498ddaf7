"""GPU parity: tcgen05 GEMM (tscd_linear) vs a plain PyTorch fp32 reference of the same op on the same
16-bit-rounded operands.  Tolerance: fp32 accumulation-order noise + one 16-bit rounding of the output."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(300, 768, 256), (128, 25, 1024), (1000, 1024, 768), (37, 4, 1024), (2500, 512, 512),
                                   (130, 1536, 256), (64, 1, 1024)])
def test_linear_matches_fp32_reference(dtype, M, N, K):
    from tscd_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + N)
    x = (torch.randn(M, K, generator=g)).to(dtype).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dtype).cuda()
    b = torch.randn(N, generator=g).cuda()
    o16, o32 = ops.linear(x, w, b, want16=True, want32=True)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    scale = ref.abs().max()
    assert (o32 - ref).abs().max() / scale < 2e-5
    tol16 = 2e-3 if dtype == torch.float16 else 1.6e-2
    assert (o16.float() - ref).abs().max() / scale < tol16


def test_linear_device_row_count_and_column_slices():
    from tscd_b200 import ops
    g = torch.Generator().manual_seed(0)
    M, N, K = 700, 512, 512
    big = torch.randn(M, 768, generator=g).half().cuda()
    x = big[:, 256:]                       # column slice: pitch 768, 16-byte aligned offset
    w = (torch.randn(N, K, generator=g) / 22).half().cuda()
    out = torch.full((M, 1024), 7.0, dtype=torch.float16, device="cuda")
    m_dev = torch.tensor([333], dtype=torch.int32, device="cuda")
    ops.linear(x, w, None, m_dev=m_dev, out16=out[:, 512:], want16=False)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    assert (out[:333, 512:].float() - ref[:333]).abs().max() / ref.abs().max() < 2e-3
    assert torch.all(out[:, :512] == 7.0)                     # neighbouring columns untouched
    assert torch.all(out[384:, 512:] == 7.0)                  # rows of CTAs past the device count untouched
