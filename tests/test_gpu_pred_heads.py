"""Conv-tower seam: the 1x1 prediction convolutions as tcgen05 GEMMs that write the fused head layout
(ops.pred_heads) against torch's conv2d on the same fp16 feature maps; the result must be consumable by K1 / K3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C", [25, 30])
def test_pred_heads_match_conv_and_feed_selection(C):
    from tscd_b200 import ops
    hw = [(24, 24), (12, 12), (6, 6)]
    Fn, K = 5, 256
    g = torch.Generator(device="cuda").manual_seed(3 + C)
    an = ops.AnchorSpec(hw)
    reg_f, cls_f, wro, bro, wc, bc, want = [], [], [], [], [], [], []
    for (h, w) in hw:
        rf = torch.randn(Fn, K, h, w, generator=g, device="cuda").half().contiguous(memory_format=torch.channels_last)
        cf = torch.randn(Fn, K, h, w, generator=g, device="cuda").half().contiguous(memory_format=torch.channels_last)
        w1 = (torch.randn(5, K, generator=g, device="cuda") / 16).half()
        w2 = (torch.randn(C, K, generator=g, device="cuda") / 16).half()
        b1 = torch.randn(5, generator=g, device="cuda") * 0.1 - 2
        b2 = torch.randn(C, generator=g, device="cuda") * 0.1 - 3
        reg_f.append(rf); cls_f.append(cf); wro.append(w1); bro.append(b1); wc.append(w2); bc.append(b2)
        ro = torch.nn.functional.conv2d(rf.float(), w1.float()[:, :, None, None], b1)          # fp32 reference on the fp16 values
        co = torch.nn.functional.conv2d(cf.float(), w2.float()[:, :, None, None], b2)
        want.append(torch.cat([ro, co], 1).flatten(2).permute(0, 2, 1))                         # [F, HW, 5+C]
    head = ops.pred_heads(reg_f, cls_f, wro, bro, wc, bc, an, C)
    torch.cuda.synchronize()
    rows, objp = head._keep
    rp = 32 if C == 25 else 64
    start = 0
    for l, (h, w) in enumerate(hw):
        got = rows[l].float()
        assert got.shape == (Fn, h * w, rp)
        err = (got[:, :, :5 + C] - want[l]).abs().max()
        assert float(err) < 2e-2, f"level {l}: {float(err)}"                                   # fp16 output rounding of logits up to ~8
        assert float(got[:, :, 5 + C:].abs().max()) == 0.0                                      # padding stays zero
        assert torch.equal(objp[:, start:start + h * w], rows[l][:, :, 4])
        start += h * w
    # the per-level fused rows feed K1 like the packed single tensor does
    flat = torch.cat([r for r in rows], 1).contiguous()
    ref_head = ops.HeadViews.from_rows(flat, objp, an, C)
    c0, c1 = ops.select(ref_head, "A", pre_k=200), ops.select(head, "A", pre_k=200)
    torch.cuda.synchronize()
    for k in ("idx", "box", "score", "cls", "count"):
        assert torch.equal(c0[k], c1[k]), k
