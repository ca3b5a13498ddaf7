"""Host side of AggregationStage.forward_host: the read-back buffers of one chunk (packed detection tables + offsets, filled on
the device by tscd_pack_rows) -> the reference's per-frame lists (`None` for frames without candidates, fresh tensors).  CPU only."""
import torch

from tscd_b200.stage import AggregationStage


def test_unpack_host_lists_and_byte_counts():
    nlf = 5
    counts = [3, 0, 2, 0, 4]                       # detections per local frame
    cand = [7, 0, 2, 5, 9]                         # candidates per frame: frame 3 had candidates but none survived the NMS
    off = [0]
    for c in counts:
        off.append(off[-1] + c)
    cap = 16
    det = torch.arange(cap * 7, dtype=torch.float32).reshape(cap, 7)
    ori = -det
    pk = dict(status=torch.zeros(1, dtype=torch.int32), sel_count=torch.tensor([30] * 8, dtype=torch.int32),
              det_cand=torch.tensor(cand + [0, 0, 0], dtype=torch.int32),
              det_offsets=torch.tensor(off + [off[-1]] * 3, dtype=torch.int32), det_packed=det,
              ori_offsets=torch.tensor(off + [off[-1]] * 3, dtype=torch.int32), ori_packed=ori, _row_bytes=3 * 256 * 2)
    res, res_ori, nbytes, zc = AggregationStage._unpack_host(pk, nlf)
    assert len(res) == nlf and len(res_ori) == nlf
    for f in range(nlf):
        if cand[f] == 0:
            assert res[f] is None and res_ori[f] is None            # post_process.py:54-55: no candidates -> None
        else:
            assert res[f].shape == (counts[f], 7)
            assert torch.equal(res[f], det[off[f]:off[f + 1]]) and torch.equal(res_ori[f], ori[off[f]:off[f + 1]])
    # fresh memory: the plan's pinned buffers are overwritten by the next replay
    res[0].zero_()
    assert float(det[0, 1]) == 1.0
    assert zc == 8 * 30 * 3 * 256 * 2
    assert nbytes == 2 * (det.numel() * 4 + 9 * 4) + (8 + 1 + 8) * 4
    # zero-copy logits add the survivors' and kept rows' logit rows
    pk2 = dict(pk, cand_count=torch.tensor([750] * 8, dtype=torch.int32), _logit_row_bytes=64)
    _, _, _, zc2 = AggregationStage._unpack_host(pk2, nlf)
    assert zc2 == zc + (8 * 750 + 8 * 30) * 64


def test_unpack_host_reports_capacity_errors():
    import pytest
    pk = dict(status=torch.tensor([-3], dtype=torch.int32))
    with pytest.raises(RuntimeError, match="capacity"):
        AggregationStage._unpack_host(pk, 1)
