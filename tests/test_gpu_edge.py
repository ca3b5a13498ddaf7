"""GPU parity of the selected-anchor WaveletsHFBlock (csrc/edge.cu through tscd_edge_patches / tscd_linear /
tscd_edge_combine) against the dense block: the oracle restatement (oracle/edge_oracle.py, pinned to the reference module by
tests/golden/edge.npz) and, when the reference package is installed on the box, the reference's own WaveletsHFBlock on CUDA.
Floating point: the kernels hold patches / sub-bands / conv outputs in the 16-bit operand type with fp32 accumulation, like the
reference's fp16 evaluation; tolerance 3e-3 (fp16) / 2e-2 (bf16) of the largest magnitude of the edge rows."""
import os
import sys

import pytest
import torch

import oracle
from oracle import edge_oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HW = [(16, 16), (8, 8), (4, 4)]


def _weights(seed, levels=3):
    g = torch.Generator().manual_seed(seed)
    w3 = [(torch.randn(256, 256, 3, 3, generator=g) / 48.0).half().float() for _ in range(levels)]
    b3 = [torch.randn(256, generator=g) * 0.1 for _ in range(levels)]
    w1 = [(torch.randn(768, 768, 1, 1, generator=g) / 28.0).half().float() for _ in range(levels)]
    b1 = [torch.randn(768, generator=g) * 0.1 for _ in range(levels)]
    return w3, b3, w1, b1


def _level_feats(Fn, dt, layout, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(3):
        lv = []
        for (h, w) in HW:
            x = torch.randn(Fn, 256, h, w, generator=g).to(dt).cuda()
            lv.append(x.contiguous() if layout == "nchw" else x.contiguous(memory_format=torch.channels_last))
        out.append(lv)
    return out


def _select(feats, edge_arg, feat_dt, bank_dt, Fn, seed=5, limits=(100, 150)):
    from tscd_b200 import ops, selection
    C = 5
    head, _ = oracle.synth_head_outputs(Fn, HW, C, dim=8, seed=seed, obj_mean=-6.0)
    decoded = oracle.decode_outputs(head, HW, [8, 16, 32])
    an = ops.AnchorSpec(HW)
    hv = ops.HeadViews.from_fused(decoded.cuda(), an, apply_sigmoid=False, apply_decode=False)
    views = (ops.view_levels(feats[0]), ops.view_levels(feats[1]), edge_arg)
    cfg = selection.SelectionConfig(mode="B", minimal_limit=limits[0], maximal_limit=limits[1], use_pre_nms=False, max_proposals=an.num_anchors)
    sel = selection.select_and_gather(hv, views, feat_dt, 256, cfg, bank_dtype=bank_dt)
    torch.cuda.synchronize()
    assert int(sel["status"].item()) == 0
    return sel


def _dense_rows(dense_levels, sel, Fn):
    """Rows of the dense per-level edge maps [F,256,H,W] at the selected anchors, bank order."""
    flat = torch.cat([d.flatten(2) for d in dense_levels], 2).permute(0, 2, 1)      # [F, A, 256]  (tscd_head.py:410-412)
    counts = sel["sel_count"].cpu().tolist()
    rows = [flat[f, sel["sel_idx"][f, :counts[f]].long().to(flat.device)] for f in range(Fn)]
    return torch.cat(rows, 0), counts


@pytest.mark.parametrize("feat_dt,bank_dt,layout", [(torch.float16, torch.float16, "cl"), (torch.float16, torch.float16, "nchw"),
                                                    (torch.float32, torch.float16, "cl"), (torch.float16, torch.bfloat16, "cl"),
                                                    (torch.bfloat16, torch.bfloat16, "cl")])
def test_selected_edge_rows_vs_dense_oracle(feat_dt, bank_dt, layout):
    from tscd_b200 import ops
    Fn = 6
    feats = _level_feats(Fn, feat_dt, layout, seed=3)
    w3, b3, w1, b1 = _weights(11)
    if bank_dt == torch.bfloat16:                       # operands are rounded to the operand type: give the oracle the same weights
        w3, w1 = [w.bfloat16().float() for w in w3], [w.bfloat16().float() for w in w1]
    block = ops.EdgeBlock(w3, b3, w1, b1, dtype=bank_dt)
    sel = _select(feats, block, feat_dt, bank_dt, Fn)
    dense = [edge_oracle.wavelets_hf_block(feats[1][l].float().cpu().contiguous(), w1[l], b1[l], w3[l], b3[l]) for l in range(3)]
    want, counts = _dense_rows(dense, sel, Fn)
    got = sel["bank_edge"][:want.shape[0]].float().cpu()
    assert want.shape[0] == sum(counts) and min(counts) >= 100
    # the selection must exercise all levels, image borders (zero padding) and all four (y&1, x&1) parities
    idx = torch.cat([sel["sel_idx"][f, :counts[f]] for f in range(Fn)]).cpu()
    assert (idx < 256).any() and ((idx >= 256) & (idx < 320)).any() and (idx >= 320).any()
    l0 = idx[idx < 256]
    y, x = l0 // 16, l0 % 16
    assert (y == 0).any() and (y == 15).any() and (x == 0).any() and (x == 15).any()
    assert len({(int(a) & 1, int(b) & 1) for a, b in zip(y.tolist(), x.tolist())}) == 4
    tol = 3e-3 if bank_dt == torch.float16 else 2e-2
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    assert scale > 0.1 and err <= tol * scale, (err, scale)
    # rows are bit-identical to a second run (slot assignment order must not matter)
    sel2 = _select(feats, block, feat_dt, bank_dt, Fn)
    assert torch.equal(sel2["bank_edge"][:want.shape[0]], sel["bank_edge"][:want.shape[0]])
    # the other two planes are untouched by the edge path
    ref = _select(feats, ops.view_levels(feats[2]), feat_dt, bank_dt, Fn)
    assert torch.equal(ref["bank_reg"][:want.shape[0]], sel["bank_reg"][:want.shape[0]])
    assert torch.equal(ref["bank_cls"][:want.shape[0]], sel["bank_cls"][:want.shape[0]])
    for f in range(Fn):
        assert torch.equal(ref["sel_rows"][f, :counts[f]], sel["sel_rows"][f, :counts[f]])


def test_selected_edge_rows_empty_and_full_frames():
    """Frames with no proposal, and every anchor of every level selected (each level's segment completely filled)."""
    from tscd_b200 import ops, selection
    Fn, C = 4, 3
    feats = _level_feats(Fn, torch.float16, "cl", seed=8)
    w3, b3, w1, b1 = _weights(12)
    block = ops.EdgeBlock(w3, b3, w1, b1, dtype=torch.float16)
    head, _ = oracle.synth_head_outputs(Fn, HW, C, dim=8, seed=2, obj_mean=[-30.0, 8.0, -30.0, 8.0])
    decoded = oracle.decode_outputs(head, HW, [8, 16, 32])
    an = ops.AnchorSpec(HW)
    hv = ops.HeadViews.from_fused(decoded.cuda(), an, apply_sigmoid=False, apply_decode=False)
    cfg = selection.SelectionConfig(mode="B", use_pre_nms=False, minimal_limit=0, maximal_limit=0, max_proposals=an.num_anchors)
    sel = selection.select_and_gather(hv, (ops.view_levels(feats[0]), ops.view_levels(feats[1]), block), torch.float16, 256, cfg,
                                      bank_dtype=torch.float16)
    torch.cuda.synchronize()
    assert int(sel["status"].item()) == 0
    counts = sel["sel_count"].cpu().tolist()
    assert counts[0] == 0 and counts[2] == 0 and counts[1] == an.num_anchors and counts[3] == an.num_anchors
    dense = [edge_oracle.wavelets_hf_block(feats[1][l].float().cpu().contiguous(), w1[l], b1[l], w3[l], b3[l]) for l in range(3)]
    want, _ = _dense_rows(dense, sel, Fn)
    got = sel["bank_edge"][:want.shape[0]].float().cpu()
    assert float((got - want).abs().max()) <= 3e-3 * float(want.abs().max())


def test_selected_edge_rows_vs_reference_module_on_cuda():
    """The reference's own WaveletsHFBlock (fp16, CUDA, dense) from the installed package vs the selected-anchor kernels."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_runner
    if not ref_runner.available():
        pytest.skip("reference package not installed (baseline/_ref)")
    ref_runner.install(cpu_redirect=False)
    from yolox.models.surrounding_extraction import WaveletsHFBlock
    from tscd_b200 import ops
    Fn = 5
    feats = _level_feats(Fn, torch.float16, "cl", seed=21)
    torch.manual_seed(7)
    mods = torch.nn.ModuleList([WaveletsHFBlock(256) for _ in range(3)]).cuda().half().eval()
    block = ops.EdgeBlock.from_modules(mods, dtype=torch.float16)
    sel = _select(feats, block, torch.float16, torch.float16, Fn, seed=9)
    with torch.no_grad():
        dense = [mods[l](feats[1][l]) for l in range(3)]
    want, counts = _dense_rows(dense, sel, Fn)
    got = sel["bank_edge"][:want.shape[0]].float()
    want = want.float()
    scale = float(want.abs().max())
    assert scale > 0 and float((got - want).abs().max()) <= 3e-3 * scale, (float((got - want).abs().max()), scale)


def test_selected_edge_rows_reject_odd_maps():
    """The reference's stride-2 Haar transform has no same-size inverse on odd feature maps (its product of x_content and x_idwt
    fails there too): the C-ABI reports the configuration as unsupported instead of producing rows."""
    from tscd_b200 import ops, selection
    hw = [(5, 6), (4, 4), (2, 2)]
    Fn, C = 2, 3
    g = torch.Generator().manual_seed(3)
    feats = [[torch.randn(Fn, 256, h, w, generator=g).half().cuda().contiguous(memory_format=torch.channels_last) for h, w in hw] for _ in range(2)]
    w3, b3, w1, b1 = _weights(5)
    block = ops.EdgeBlock(w3, b3, w1, b1, dtype=torch.float16)
    head, _ = oracle.synth_head_outputs(Fn, hw, C, dim=8, seed=4, obj_mean=2.0)
    decoded = oracle.decode_outputs(head, hw, [8, 16, 32])
    an = ops.AnchorSpec(hw)
    hv = ops.HeadViews.from_fused(decoded.cuda(), an, apply_sigmoid=False, apply_decode=False)
    cfg = selection.SelectionConfig(mode="B", use_pre_nms=False, minimal_limit=4, maximal_limit=8, max_proposals=an.num_anchors)
    with pytest.raises(RuntimeError, match="unsupported"):
        selection.select_and_gather(hv, (ops.view_levels(feats[0]), ops.view_levels(feats[1]), block), torch.float16, 256, cfg,
                                    bank_dtype=torch.float16)
    torch.cuda.synchronize()
