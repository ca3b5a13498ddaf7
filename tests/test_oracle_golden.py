"""Pin oracle/ against outputs of the reference's own PyTorch code (tests/golden/*.npz, produced by
tools/make_goldens.py in the build container).  Indices must match exactly; floats to fp32 noise."""
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN, unpack_list


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _sd(z):
    return {k[3:]: _t(z[k]) for k in z.files if k.startswith("sd.")}


def _close(a, b, tol=2e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size:
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-6)
        assert err <= tol, err


@pytest.fixture(scope="module")
def sel():
    return np.load(os.path.join(GOLDEN, "select.npz"))


def test_decode_outputs(sel):
    hw = [tuple(x) for x in sel["hw"].tolist()]
    dec = oracle.decode_outputs(_t(sel["head_out"]), hw, sel["strides"].tolist())
    _close(dec, sel["decoded"], 1e-6)


@pytest.mark.parametrize("name,kw", [
    ("b_min50_max500", dict(minimal_limit=50, maximal_limit=500, use_pre_nms=False)),
    ("b_max100", dict(minimal_limit=0, maximal_limit=100, use_pre_nms=False)),
    ("b_prenms", dict(minimal_limit=0, maximal_limit=0, use_pre_nms=True)),
    ("b_min700_prenms", dict(minimal_limit=700, maximal_limit=0, use_pre_nms=True)),
])
def test_select_mode_b(sel, name, kw):
    rows, idxs = oracle.select_mode_b(_t(sel["decoded"]), int(sel["C"]), nms_thre=0.75, **kw)
    g_rows, g_idx = unpack_list(sel, name + ".rows"), unpack_list(sel, name + ".idx")
    for f in range(len(g_rows)):
        assert idxs[f].tolist() == g_idx[f].tolist()
        assert np.array_equal(rows[f].numpy(), g_rows[f])


def test_select_mode_a(sel):
    rows, idxs = oracle.select_mode_a(_t(sel["decoded"]), int(sel["C"]), nms_thre=0.75, pre_k=750, top_k=30)
    g_rows, g_idx = unpack_list(sel, "a_750_30.rows"), unpack_list(sel, "a_750_30.idx")
    for f in range(len(g_rows)):
        assert idxs[f].tolist() == g_idx[f].tolist()
        assert np.array_equal(rows[f].numpy(), g_rows[f])


def test_msa_yolov():
    z = np.load(os.path.join(GOLDEN, "msa.npz"))
    sd = _sd(z)
    x_c, x_r, r2c, r2o = oracle.attention_msa(sd, "trans.msa.", _t(z["x_cls"]), _t(z["x_reg"]), _t(z["cls_score"]))
    _close(x_c, z["att_x_cls"]); _close(x_r, z["att_x_reg"]); _close(r2c, z["r2c"]); _close(r2o, z["r2o"])
    assert np.array_equal(r2c.numpy() > 0, z["r2c"] > 0)
    out, _ = oracle.msa_yolov(sd, "trans.", _t(z["x_cls"]), _t(z["x_reg"]), _t(z["cls_score"]))
    _close(out, z["out_cls"])


@pytest.fixture(scope="module")
def stage():
    return np.load(os.path.join(GOLDEN, "stage_tscd.npz"))


def test_timing_signal(stage):
    te = oracle.timing_signal_1d(torch.arange(3, 6), 256)
    _close(te, stage["clip1.time_embedding"], 1e-6)


def test_stage_tscd_two_clips_with_resume(stage):
    z = stage
    sd = _sd(z)
    C, L, G = int(z["C"]), int(z["L"]), int(z["G"])
    hw = [tuple(x) for x in z["hw"].tolist()]
    state = None
    for clip in range(2):
        p = f"clip{clip}."
        decoded = oracle.decode_outputs(_t(z[p + "head_out"]), hw, z["strides"].tolist())
        _close(decoded, z[p + "decoded"], 1e-6)
        trace = {}
        res, res_ori, state = oracle.stage_tscd(
            sd, decoded, _t(z[p + "plane_cls"]), _t(z[p + "plane_reg"]), _t(z[p + "plane_edge"]),
            _t(z[p + "time_embedding"]), C, L, G, selection="B",
            select_kwargs=dict(nms_thre=0.75, minimal_limit=12, maximal_limit=40, use_pre_nms=False),
            nms_thresh=0.5, resume=(clip == 1), state=state, trace=trace)
        g_idx = unpack_list(z, p + "idx")
        for f in range(L + G):
            assert trace["idxs"][f].tolist() == g_idx[f].tolist()
        for j, nm in enumerate(("bank_cls", "bank_reg", "bank_edge", "cls_scores", "fg_scores", "locs", "all_scores")):
            assert np.array_equal(trace["bank"][j].numpy(), z[p + nm]), nm
        _close(trace["agg_cls"], z[p + "agg_cls"])
        _close(trace["iou_cls"], z[p + "iou_cls"])
        _close(trace["iou_reg"], z[p + "iou_reg"])
        # Hungarian: same costs, same assignment
        g_cost, g_col = unpack_list(z, p + "lap_cost"), unpack_list(z, p + "lap_col")
        assert len(g_col) == len(trace["cafm"]["perm"])
        for k in range(len(g_col)):
            _close(trace["cafm"]["cost"][k], g_cost[k], 1e-5)
            assert trace["cafm"]["perm"][k][:len(g_col[k])].tolist() == g_col[k].tolist()
        _close(trace["matched"], z[p + "fc_reg_matcher"], 5e-5)
        _close(trace["obj_ref"], z[p + "task_aligned"], 5e-5)
        _close(trace["cls_preds"], z[p + "cls_preds"], 5e-5)
        _close(trace["obj_preds"], z[p + "obj_preds"], 5e-5)
        _close(trace["reg_deltas"], z[p + "reg_deltas"], 5e-5)
        g_res, g_ori = unpack_list(z, p + "result"), unpack_list(z, p + "result_ori")
        for f in range(L):
            for got, want in ((res[f], g_res[f]), (res_ori[f], g_ori[f])):
                if want is None:
                    assert got is None
                    continue
                assert got.shape == want.shape
                assert np.array_equal(got[:, 6].numpy(), want[:, 6])        # same classes, same order
                _close(got, want, 5e-5)
