"""Pin the plain-C restatements of torchvision nms / scipy linear_sum_assignment (oracle/oracle_kernels.c)
against (a) the committed known-answer vectors and (b) the installed binaries, black-box."""
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLDEN, "thirdparty.npz"))


def test_nms_known_answers(kat):
    for t in range(5):
        boxes = torch.from_numpy(kat[f"nms{t}.boxes"])
        scores = torch.from_numpy(kat[f"nms{t}.scores"])
        cls = torch.from_numpy(kat[f"nms{t}.cls"])
        for thr in (0.5, 0.75):
            keep = oracle.batched_nms(boxes, scores, cls, thr)
            assert keep.tolist() == kat[f"nms{t}.{thr}.keep"].tolist()


def test_lap_known_answers(kat):
    for t in range(8):
        r, c = oracle.lap(kat[f"lap{t}.cost"])
        assert r.tolist() == kat[f"lap{t}.row"].tolist()
        assert c.tolist() == kat[f"lap{t}.col"].tolist()


def test_nms_blackbox_vs_torchvision():
    tv = pytest.importorskip("torchvision")
    from torchvision.ops import boxes as tvb
    g = torch.Generator().manual_seed(123)
    for trial in range(60):
        n = int(torch.randint(1, 400, (1,), generator=g))
        ctr = torch.rand(max(n // 8, 1), 2, generator=g) * 300 - 30
        c = ctr[torch.randint(0, ctr.shape[0], (n,), generator=g)] + torch.randn(n, 2, generator=g) * 3
        wh = torch.rand(n, 2, generator=g) * 30 + 10
        boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
        scores = torch.rand(n, generator=g)
        scores[torch.randint(0, n, (n // 4 + 1,), generator=g)] = 0.5     # many exact ties
        cls = torch.randint(0, 5, (n,), generator=g).float()
        thr = (0.3, 0.5, 0.75)[trial % 3]
        ref = tvb._batched_nms_coordinate_trick(boxes, scores, cls, thr)
        got = oracle.batched_nms(boxes, scores, cls, thr)
        assert got.tolist() == ref.tolist(), f"trial {trial}"
        assert oracle.nms(boxes, scores, thr).tolist() == tv.ops.nms(boxes, scores, thr).tolist()


def test_nms_empty_and_single():
    assert oracle.batched_nms(torch.zeros(0, 4), torch.zeros(0), torch.zeros(0), 0.5).numel() == 0
    assert oracle.nms(torch.tensor([[0., 0., 1., 1.]]), torch.tensor([0.3]), 0.5).tolist() == [0]


def test_lap_blackbox_vs_scipy():
    sp = pytest.importorskip("scipy.optimize")
    rng = np.random.default_rng(7)
    for trial in range(80):
        r, c = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        cost = rng.random((r, c))
        if trial % 4 == 0:
            cost = np.round(cost * 4) / 4            # heavy ties
        if trial % 7 == 0:
            cost = cost.astype(np.float32).astype(np.float64)
        a, b = sp.linear_sum_assignment(cost)
        ga, gb = oracle.lap(cost)
        assert ga.tolist() == a.tolist() and gb.tolist() == b.tolist(), f"trial {trial} shape {(r, c)}"
