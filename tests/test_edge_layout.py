"""Host-side layout of the selected-anchor wavelet block (tscd_b200.ops.EdgeBlock): the tap-major 3x3 weight matrix, the
[LH|HL|HH] sub-band order and the inverse-Haar signs the kernels use (csrc/edge.cu), checked WITHOUT a GPU by evaluating the
same per-anchor formula with torch matmuls on the CPU against the dense oracle block (oracle/edge_oracle.py, itself pinned to
the reference module by tests/golden/edge.npz)."""
import torch
import torch.nn.functional as F

from oracle import edge_oracle
from tscd_b200 import ops


def _per_anchor(block, x, l):
    """What tscd_edge_patches + two GEMMs + tscd_edge_combine compute, for EVERY anchor of one level (fp32 on the CPU)."""
    Fn, C, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1))
    taps = [xp[:, :, dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)]          # tap-major: (ky, kx), then channel
    patches = torch.stack(taps, 1).permute(0, 3, 4, 1, 2).reshape(Fn * H * W, 9 * C)
    content = torch.relu(patches @ block.w3[l].float().t() + block.b3[l])
    a, b = x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2]
    c, d = x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2]
    hf = torch.cat([0.5 * (a + b - c - d), 0.5 * (a - b + c - d), 0.5 * (a - b - c + d)], 1)      # [LH | HL | HH]
    hf = hf.repeat_interleave(2, 2).repeat_interleave(2, 3)                                # the anchor's 2x2 block
    hf = hf.permute(0, 2, 3, 1).reshape(Fn * H * W, 3 * C)
    hfo = torch.relu(hf @ block.w1[l].float().t() + block.b1[l])
    yy = torch.arange(H).view(1, H, 1).expand(Fn, H, W).reshape(-1)
    xx = torch.arange(W).view(1, 1, W).expand(Fn, H, W).reshape(-1)
    s_lh = torch.where(yy % 2 == 1, -1.0, 1.0).unsqueeze(1)
    s_hl = torch.where(xx % 2 == 1, -1.0, 1.0).unsqueeze(1)
    idwt = 0.5 * (s_lh * hfo[:, :C] + s_hl * hfo[:, C:2 * C] + s_lh * s_hl * hfo[:, 2 * C:])
    return (content * idwt).reshape(Fn, H, W, C).permute(0, 3, 1, 2)


def test_edge_block_layout_matches_dense_oracle():
    torch.manual_seed(0)
    C = 256
    w3 = [torch.randn(C, C, 3, 3) / 48 for _ in range(2)]
    b3 = [torch.randn(C) * 0.1 for _ in range(2)]
    w1 = [torch.randn(3 * C, 3 * C, 1, 1) / 28 for _ in range(2)]
    b1 = [torch.randn(3 * C) * 0.1 for _ in range(2)]
    block = ops.EdgeBlock(w3, b3, w1, b1, dtype=torch.float16, device="cpu")
    assert block.w3[0].shape == (C, 9 * C) and block.w1[0].shape == (3 * C, 3 * C) and block.b3[0].dtype == torch.float32
    for l, (H, W) in enumerate([(4, 6), (2, 2)]):
        x = torch.randn(2, C, H, W)
        want = edge_oracle.wavelets_hf_block(x, block.w1[l].float().reshape(3 * C, 3 * C, 1, 1), block.b1[l],
                                             block.w3[l].float().reshape(C, 3, 3, C).permute(0, 3, 1, 2).contiguous(), block.b3[l])
        got = _per_anchor(block, x, l)
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), float((got - want).abs().max())
