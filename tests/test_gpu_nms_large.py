"""GPU parity of the workspace NMS path (csrc/nms_large.cu, cand_cap > 4096) against the oracle's restatement of
torchvision.ops.batched_nms: keep lists bit-exact (order included).

The shape that matters is the final per-class NMS of post_process.py:36-65 for the shipped OVIS-L limits
(exps/TSCD_OVIS/ovis_tscd_large.py:45,49): up to 500 proposals x 25 classes = 12 500 (proposal, class) rows per frame,
every class holding the SAME boxes with different scores."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _check(box, score, cls, count, thr, max_keep=None, label=""):
    from tscd_b200 import ops
    keep, kc, status = ops.nms(box.cuda(), score.cuda(), cls.cuda(), count.cuda(), thr, max_keep=max_keep)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    supp = 0
    for f in range(box.shape[0]):
        n = int(count[f])
        want = oracle.batched_nms(box[f, :n], score[f, :n], cls[f, :n].float(), thr).tolist()
        supp += n - len(want)
        if max_keep is not None:
            want = want[:max_keep]
        got = keep[f, :int(kc[f])].cpu().tolist()
        assert got == want, f"{label} frame {f} n {n}: first diff at {next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want)))} ({len(got)} vs {len(want)})"
    return supp


def _expanded_frame(g, n_prop, C, spread, neg=0.0, dup_scores=False):
    """(proposal, class) expansion like post_process.py:36-46: every class sees the same boxes."""
    ctr = torch.rand(max(n_prop // 6, 1), 2, generator=g) * spread - neg
    which = torch.randint(0, ctr.shape[0], (n_prop,), generator=g)
    c = ctr[which] + torch.randn(n_prop, 2, generator=g) * 4
    wh = (torch.rand(ctr.shape[0], 2, generator=g) * 60 + 30)[which] * (1 + 0.08 * torch.randn(n_prop, 2, generator=g))
    pb = torch.cat([c - wh / 2, c + wh / 2], 1)
    present = torch.rand(n_prop, C, generator=g) < 0.97            # a few (proposal, class) pairs fall below the 0.001 filter
    p_idx, c_idx = torch.where(present)                              # row-major (proposal, class) order
    sc = torch.rand(len(p_idx), generator=g)
    if dup_scores:
        sc = (sc * 64).round() / 64                                  # heavy exact ties, also across classes
    return pb[p_idx], sc, c_idx.int()


def test_final_nms_shape_500x25_dense():
    """12 500-row frames (500 proposals x 25 classes), ragged neighbours (small / empty frames in the same launch), with and
    without exact score ties, both IoU thresholds the stage uses, plus max_keep truncation."""
    g = torch.Generator().manual_seed(7)
    C, cap = 25, 500 * 25
    specs = [(500, False), (500, True), (0, False), (37, False), (300, True), (499, False)]
    Fn = len(specs)
    box = torch.zeros(Fn, cap, 4)
    score = torch.zeros(Fn, cap)
    cls = torch.zeros(Fn, cap, dtype=torch.int32)
    count = torch.zeros(Fn, dtype=torch.int32)
    for f, (n_prop, dup) in enumerate(specs):
        if n_prop == 0:
            continue
        b, s, c = _expanded_frame(g, n_prop, C, spread=520.0, dup_scores=dup)
        n = len(s)
        box[f, :n], score[f, :n], cls[f, :n], count[f] = b, s, c, n
    assert int(count.max()) > 11000
    for thr in (0.5, 0.75):
        supp = _check(box, score, cls, count, thr, label=f"thr {thr}")
        assert supp > 5000                                          # real suppression chains
    _check(box, score, cls, count, 0.5, max_keep=1000, label="max_keep 1000")


def test_cross_class_interaction_takes_the_exact_general_path():
    """Negative coordinates make the class x-bands of the coordinate trick overlap; a box of class c+1 at (-11,-11,-1,-1)
    lands EXACTLY on a class-c box at (M-10, M-10, M, M) after the offset (M = boxes.max()), so a cross-class pair really
    suppresses: the per-class decomposition is invalid and the frame must be redone by the general algorithm.  A second
    frame only has overlapping bands (no cross-class hit) and stays on the per-class path; both must match the oracle."""
    g = torch.Generator().manual_seed(9)
    C, n_prop = 12, 450
    cap = 6000
    box = torch.zeros(2, cap, 4)
    score = torch.zeros(2, cap)
    cls = torch.zeros(2, cap, dtype=torch.int32)
    count = torch.zeros(2, dtype=torch.int32)
    for f in range(2):
        b, s, c = _expanded_frame(g, n_prop, C, spread=400.0, neg=60.0)
        n = len(s)
        if f == 0:
            M = 600.0
            extra_b = torch.tensor([[M - 10, M - 10, M, M], [-11.0, -11.0, -1.0, -1.0], [M - 10, M - 10, M, M]])
            extra_s = torch.tensor([0.99, 0.98, 0.97])               # class 3 box kept, class 4 twin suppressed ACROSS classes, class 5?
            extra_c = torch.tensor([3, 4, 2], dtype=torch.int32)
            b, s, c = torch.cat([b, extra_b]), torch.cat([s, extra_s]), torch.cat([c, extra_c])
            n += 3
            assert float(b.max()) == M
        box[f, :n], score[f, :n], cls[f, :n], count[f] = b, s, c, n
    # the engineered pair is really a cross-class suppression in the reference semantics
    n0 = int(count[0])
    want = oracle.batched_nms(box[0, :n0], score[0, :n0], cls[0, :n0].float(), 0.5).tolist()
    assert (n0 - 3) in want and (n0 - 2) not in want
    _check(box, score, cls, count, 0.5, label="cross-class")


def test_one_huge_class_and_wide_class_ids():
    """A class with more members than one CTA resolves (> 2048) and class ids outside [0, 256) both fall back to the general
    path; mixed with a normal frame in the same launch."""
    g = torch.Generator().manual_seed(13)
    cap = 5000
    Fn = 3
    box = torch.zeros(Fn, cap, 4)
    score = torch.rand(Fn, cap, generator=g)
    cls = torch.zeros(Fn, cap, dtype=torch.int32)
    count = torch.tensor([4800, 4500, 5000], dtype=torch.int32)
    for f in range(Fn):
        n = int(count[f])
        c = torch.rand(n, 2, generator=g) * 700
        wh = torch.rand(n, 2, generator=g) * 50 + 15
        box[f, :n] = torch.cat([c - wh / 2, c + wh / 2], 1)
    cls[0, :4800] = 7                                                # one class of 4800 members
    cls[0, ::9] = 1
    cls[1, :4500] = torch.randint(0, 400, (4500,), generator=g).int()    # ids up to 399
    cls[2] = torch.randint(0, 30, (cap,), generator=g).int()
    _check(box, score, cls, count, 0.5, label="fallbacks")
    _check(box, score, cls, count, 0.75, max_keep=64, label="fallbacks top-64")
