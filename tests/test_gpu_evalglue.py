"""GPU parity of the evaluator / Predictor glue against the REFERENCE's own consumers run on the same
detections: OVISEvaluator.convert_to_coco_format (yolox/evaluators/ovis_evaluator_v2.py:233-289) and Predictor.to_repp_heavy
(tools/val_to_imdb.py:193-218).  fp32 detections -> the dicts must be identical (keys, order, values bit for bit)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_runner  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_runner.available(), reason="reference package not installed (baseline/_ref)")]


def _detections(Fn, cap, seed):
    g = torch.Generator().manual_seed(seed)
    counts = [0, 1, cap, 17, 0, 33, 5, cap - 1][:Fn]
    rows = torch.zeros(Fn, cap, 7)
    for f, n in enumerate(counts):
        xy = torch.rand(n, 2, generator=g) * 500 - 20
        wh = torch.rand(n, 2, generator=g) * 200 + 1
        rows[f, :n, :2], rows[f, :n, 2:4] = xy, xy + wh
        rows[f, :n, 4:6] = torch.rand(n, 2, generator=g)
        rows[f, :n, 6] = torch.randint(0, 25, (n,), generator=g).float()
    return rows, counts


def test_coco_and_repp_dicts_match_the_reference_consumers():
    from tscd_b200 import evalglue
    ref_runner.install(cpu_redirect=False)
    from yolox.evaluators.ovis_evaluator_v2 import OVISEvaluator
    Fn, cap = 8, 40
    rows, counts = _detections(Fn, cap, 3)
    info = [(360 + 40 * f, 640 - 30 * f) for f in range(Fn)]             # (height, width) of every frame
    img_size = (576, 576)
    scales = [min(img_size[0] / float(h), img_size[1] / float(w)) for h, w in info]
    outputs = [None if n == 0 else rows[f, :n].clone() for f, n in enumerate(counts)]
    # ---- reference: convert_to_coco_format (no dataloader needed: the method only reads img_size / ids) ----
    ev = OVISEvaluator(None, img_size, 0.001, 0.5, 25)
    ev.id = 100
    labels = [torch.zeros(0, 5) for _ in range(Fn)]
    want, _ = ev.convert_to_coco_format(copy.deepcopy(outputs), info, labels)
    # ---- ours: one pack kernel, one D2H ----
    cnt = torch.tensor(counts, dtype=torch.int32).cuda()
    packed, offsets = evalglue.pack_detections(rows.cuda().contiguous(), cnt, torch.tensor(scales))
    table, off = evalglue.to_host(packed, offsets)
    assert off == [0] + np.cumsum(counts).tolist()
    got = evalglue.coco_predictions(table.tolist(), off, counts, first_image_id=100)
    assert len(got) == len(want) == sum(counts)
    for a, b in zip(got, want):
        assert a == b, (a, b)
    # ---- Predictor.to_repp_heavy ----
    sys.path.insert(0, os.path.join(ref_tools(), ""))
    from val_to_imdb import Predictor
    pred = Predictor.__new__(Predictor)
    ratios = [0.9 + 0.05 * f for f in range(Fn)]
    packed_r, offsets_r = evalglue.pack_detections(rows.cuda().contiguous(), cnt, torch.tensor(ratios))
    table_r, off_r = evalglue.to_host(packed_r, offsets_r)
    got_r = evalglue.repp_predictions(table_r, off_r, counts, info, [f"img{f}" for f in range(Fn)])
    for f in range(Fn):
        want_r = pred.to_repp_heavy(None if counts[f] == 0 else rows[f, :counts[f]].clone(), ratios[f], info[f], f"img{f}")
        assert len(got_r[f]) == len(want_r), f
        for a, b in zip(got_r[f], want_r):
            assert a["image_id"] == b["image_id"]
            assert [float(x) for x in a["bbox"]] == [float(x) for x in b["bbox"]], (a, b)
            assert [float(x) for x in a["bbox_center"]] == [float(x) for x in b["bbox_center"]]
            assert [float(x) for x in a["scores"]] == [float(x) for x in b["scores"]]


def ref_tools():
    import ref_shim
    root = ref_shim.reference_root()
    for cand in (os.path.join(root, "tools"), os.path.join(root, "yolox", "tools")):
        if os.path.exists(os.path.join(cand, "val_to_imdb.py")):
            return cand
    pytest.skip("tools/val_to_imdb.py is not part of the installed reference package")
