"""Parity measurement helpers shared by the GPU stage tests.

The stage mixes float tensors (fp16 tensor-core operands, fp32 accumulation) with DISCRETE decisions taken on them
(Hungarian assignment, 0.001 score filters, IoU > thr suppression).  "Green" parity therefore means three things, all
asserted by the tests that use this module:

  1. every float tensor is within a stated tolerance of the fp32 oracle;
  2. every discrete stage is EXACT on the inputs it actually saw (device LSAP == scipy on the device's own cost table;
     device final detections == the oracle's post-processing run on the device's own logits / deltas);
  3. wherever a discrete result still differs from the all-fp32 oracle, the ORACLE's own decision margin at that point is
     below a stated epsilon (cost gap of the two assignments, |score - 0.001|, |IoU - thr|), i.e. the flip is a measured
     near-tie and not an unexplained difference.
"""
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

# stated tolerances -------------------------------------------------------------------------------------------------
TOL_NORM = 1e-3        # max |a-b| / max|b|  per tensor (fp16 operands; measured 2e-4 .. 6.2e-4 over all stage tests)
TOL_REL = 1e-2         # north_star: max relative error, elementwise, over elements with |b| >= 10 % of max|b| (measured <= 5e-3)
COST_DELTA = 2e-4      # max |device cost - oracle cost| of one matching-table entry (cosine of 1024-dim fp16-operand GEMM rows; measured 7.7e-5)
EPS_SCORE = 1e-2       # relative distance of a score to the 0.001 filter below which the filter decision may flip
EPS_IOU = 5e-3         # distance of an IoU to the NMS threshold below which the suppression decision may flip
TOL_BOX_PX = 0.25      # refined boxes (deltas x box size up to ~500 px, fp16-operand regression head)
TOL_SCORE = 1e-2       # relative, obj / class scores of matched detections


def float_err(a, b):
    """(max-normalised error, max elementwise relative error over the elements that matter)."""
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    mx = float(b.abs().max().clamp_min(1e-6))
    floor = 0.10 * mx
    return float(d.max() / mx), float((d / b.abs().clamp_min(floor)).max())


def lap_pairs(col):
    """{(row, col)} of an assignment given col4row with -1 = unmatched."""
    return {(r, int(c)) for r, c in enumerate(col) if c >= 0}


def check_device_lap_exact(cost_dev, n_ref, n_cur, lap_col_dev):
    """Discrete stage exact on its own inputs: the device assignment must equal scipy's on the DEVICE cost table."""
    ri, ci = linear_sum_assignment(cost_dev[:n_ref, :n_cur].astype(np.float64))
    want = {(int(r), int(c)) for r, c in zip(ri, ci)}
    got = lap_pairs(lap_col_dev[:n_ref])
    return got == want


def assignment_gap(cost_ora, row_map, dev_pairs, ora_pairs):
    """Cost of the device's assignment minus the oracle's, both evaluated in the ORACLE's cost table.
    row_map[r_dev] = oracle row index of device reference row r_dev.  Returns (gap, number of differing pairs)."""
    dev_in_ora = {(int(row_map[r]), c) for r, c in dev_pairs}
    tot_d = sum(float(cost_ora[r, c]) for r, c in dev_in_ora)
    tot_o = sum(float(cost_ora[r, c]) for r, c in ora_pairs)
    return tot_d - tot_o, len(dev_in_ora ^ ora_pairs) // 2 + len(dev_in_ora ^ ora_pairs) % 2


def iou_matrix(a, b):
    """Plain fp64 IoU of xyxy boxes [n,4] x [m,4]."""
    a, b = a.double(), b.double()
    lt = torch.maximum(a[:, None, :2], b[None, :, :2])
    rb = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp_min(0)
    inter = wh[..., 0] * wh[..., 1]
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter)


def explain_detection_diffs(cand, keep, det_dev, thr, conf=0.001, label="", eps_score=None, eps_iou=None, tol_box=None, tol_score=None):
    """Compare the device's final detections of one frame with the oracle's.

    cand [m,7]: the oracle's candidate rows fed to batched_nms (box4, obj, cls score, class); keep: its keep list (order =
    descending score); det_dev [k,7]: the device's detections in its keep order.
    Every device detection is mapped to the oracle candidate it is (same class, box within TOL_BOX_PX, scores within
    TOL_SCORE); detections present on one side only must be explained by a measured near-tie in the ORACLE's numbers:
      - its score is within EPS_SCORE of the 0.001 filter, or
      - an IoU between it and a same-class candidate is within EPS_IOU of `thr`, or
      - it overlaps (> thr) an already explained flip of the same class (cascade), or
      - it swaps order with a same-class candidate whose score is within EPS_SCORE (relative) and overlaps it.
    Returns (n_common, n_explained).  Raises AssertionError on an unexplained difference."""
    EPS_SCORE_, EPS_IOU_ = (EPS_SCORE if eps_score is None else eps_score), (EPS_IOU if eps_iou is None else eps_iou)
    TOL_BOX_, TOL_SCORE_ = (TOL_BOX_PX if tol_box is None else tol_box), (TOL_SCORE if tol_score is None else tol_score)
    cand = cand.float().cpu()
    det_dev = det_dev.float().cpu()
    score = cand[:, 4] * cand[:, 5]
    mapped, dev_only = [], []
    used = set()
    for j in range(det_dev.shape[0]):
        same = torch.where(cand[:, 6] == det_dev[j, 6])[0]
        best, best_d = -1, None
        if len(same):
            d = (cand[same, :4] - det_dev[j, :4]).abs().max(dim=1).values
            for t in torch.argsort(d).tolist():
                i = int(same[t])
                if i in used:
                    continue
                if float(d[t]) <= TOL_BOX_ and abs(float(cand[i, 4] - det_dev[j, 4])) <= TOL_SCORE_ * max(float(cand[i, 4]), 1e-3) + 1e-5 \
                        and abs(float(cand[i, 5] - det_dev[j, 5])) <= TOL_SCORE_ * max(float(cand[i, 5]), 1e-3) + 1e-5:
                    best, best_d = i, float(d[t])
                break
        if best >= 0:
            used.add(best)
            mapped.append(best)
        else:
            dev_only.append(j)
            mapped.append(-1)
    # a device detection with no oracle candidate passed the 0.001 filters on the device only: must be a filter near-tie
    for j in dev_only:
        s = float(det_dev[j, 4] * det_dev[j, 5])
        near = abs(s - conf) <= EPS_SCORE_ * conf or abs(float(det_dev[j, 5]) - conf) <= EPS_SCORE_ * conf
        assert near, f"{label}: device detection {j} (class {int(det_dev[j, 6])}, score {s:.3e}) has no oracle candidate and is not a filter near-tie"
    D = {i for i in mapped if i >= 0}
    O = set(int(i) for i in keep.tolist())
    diff = sorted(D ^ O, key=lambda i: (-float(score[i]), i))
    explained = set()
    for x in diff:
        same = torch.where(cand[:, 6] == cand[x, 6])[0]
        same = same[same != x]
        sx = float(score[x])
        ok = abs(sx - conf) <= EPS_SCORE_ * conf or abs(float(cand[x, 5]) - conf) <= EPS_SCORE_ * conf
        if not ok and len(same):
            ix = iou_matrix(cand[x:x + 1, :4], cand[same, :4])[0]
            ok = bool(((ix - thr).abs() <= EPS_IOU_).any())
            if not ok:
                ok = any(int(y) in explained and float(ix[t]) > thr for t, y in enumerate(same.tolist()))
            if not ok:
                close = ((score[same] - sx).abs() <= EPS_SCORE_ * sx) & (ix > thr)
                ok = bool(close.any())
        assert ok, (f"{label}: candidate {x} (class {int(cand[x, 6])}, score {sx:.4e}) is {'kept' if x in O else 'dropped'} by the "
                    f"oracle but not by the device, and no oracle decision margin is near its threshold")
        explained.add(x)
    # order of the common detections: identical up to swaps of near-equal scores
    common_dev = [i for i in mapped if i in O]
    common_ora = [int(i) for i in keep.tolist() if int(i) in D]
    for a, b in zip(common_dev, common_ora):
        if a != b:
            assert abs(float(score[a] - score[b])) <= EPS_SCORE_ * float(score[b]), f"{label}: order differs beyond a score near-tie"
    return len(common_ora), len(diff) + len(dev_only)
