"""GPU parity of K5 (CAFM) in isolation: the chain kernel gets the ORACLE's fp32 matching embeddings, so the
cost matrices agree to fp32 rounding and the device LSAP must reproduce scipy's assignment exactly, for ragged
frames (n_prev <, ==, > n_cur, empty frames) and with resume across calls."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("counts_calls,kmax", [
    ([[[18, 18, 24], [30, 12, 30]], [[16, 29, 17], [5, 40, 0]]], 40),   # two clips, two consecutive calls (generic chain)
    ([[[0, 9, 33], [1, 1, 2]]], 40),                                      # empty first frame, tiny frames
    # kmax <= 32: the shared-memory / mma.sync chain (16-bit tensor-core operands inside the recurrence)
    ([[[18, 18, 24], [30, 12, 30]], [[16, 29, 17], [5, 32, 0]]], 32),
    ([[[0, 9, 31, 0, 16, 17], [1, 1, 2, 32, 32, 3]], [[0, 0, 0, 0, 0, 0], [7, 0, 30, 30, 0, 1]], [[4, 4, 4, 4, 4, 4], [0, 3, 0, 3, 0, 3]]], 32),
])
def test_cafm_chain_exact_assignments(counts_calls, kmax):
    from tscd_b200 import aggregate, ops, selection, stage
    dtype, D, C = torch.float16, 256, 5
    sd = oracle.init_stage_weights(C, dim=D, seed=23)
    sd16 = {k: (v.to(dtype).float() if v.dim() == 2 and "CA.fc" not in k else v.clone()) for k, v in sd.items()}
    cfg = stage.StageConfig(num_classes=C, selection=selection.SelectionConfig(mode="B", maximal_limit=40), dtype=dtype)
    st = stage.AggregationStage(cfg, sd)
    B, Lf = len(counts_calls[0]), len(counts_calls[0][0])
    F = Lf + 1
    state = stage.CAFMState(B, kmax, D)
    o_state = [None] * B
    g = torch.Generator().manual_seed(4)
    for call, counts in enumerate(counts_calls):
        per_frame = []
        for b in range(B):
            per_frame += counts[b] + [3]                                 # one global frame per clip (unused by CAFM)
        N = sum(per_frame)
        row_cap = ((N + 127) // 128 + 1) * 128
        loc_total = sum(sum(c) for c in counts)
        loc_cap = max(128, ((loc_total + 127) // 128) * 128)
        cnt = torch.tensor(per_frame, dtype=torch.int32).cuda()
        lay = aggregate.make_layout(cnt, B, F, Lf, row_cap, loc_cap, 128 * ((F * kmax + 127) // 128), dtype)
        feat = torch.randn(N, D, generator=g).to(dtype)
        edge = torch.randn(N, D, generator=g).to(dtype)
        base = torch.randn(12, 4 * D, generator=g)                      # shared objects -> meaningful matching
        emb_r = torch.zeros(loc_cap, 4 * D)
        emb_c = torch.zeros(loc_cap, 4 * D)
        emb_r[:loc_total] = base[torch.randint(0, 12, (loc_total,), generator=g)] + 0.5 * torch.randn(loc_total, 4 * D, generator=g)
        emb_c[:loc_total] = base[torch.randint(0, 12, (loc_total,), generator=g)] + 0.5 * torch.randn(loc_total, 4 * D, generator=g)
        if kmax <= 32:       # frames of <= 32 proposals match on the 16-bit GEMM outputs (tensor-core cost kernel): same values on both sides
            emb_r, emb_c = emb_r.to(dtype).float(), emb_c.to(dtype).float()
        te = torch.cat([oracle.timing_signal_1d(torch.arange(call * Lf, (call + 1) * Lf), 256) for _ in range(B)], 0)
        bank_reg = torch.zeros(row_cap, D, dtype=dtype).cuda(); bank_reg[:N] = feat.cuda()
        bank_edge = torch.zeros(row_cap, D, dtype=dtype).cuda(); bank_edge[:N] = edge.cuda()
        status = torch.zeros(1, dtype=torch.int32).cuda()
        resume = torch.full((B,), int(call > 0), dtype=torch.int32).cuda()
        c16, c32, perm, _ = st.run_cafm(lay, bank_reg, bank_edge, emb_r.cuda(), emb_c.cuda(), te, kmax, state, resume, status,
                                        want_debug=True)
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        off = np.concatenate([[0], np.cumsum(per_frame)])
        lpos = 0
        for b in range(B):
            rows = np.concatenate([np.arange(off[b * F + f], off[b * F + f + 1]) for f in range(Lf)]).astype(np.int64)
            nl = len(rows)
            dbg = {}
            want, o_state[b] = oracle.aware_position_reg_matcher(
                sd16, "local_reg_matcher.", feat[rows].float(), emb_r[lpos:lpos + nl], emb_c[lpos:lpos + nl], edge[rows].float(),
                counts[b], te.to(dtype).float()[b * Lf:(b + 1) * Lf], resume=(call > 0), state=o_state[b], debug=dbg)
            o_perm = np.concatenate(dbg["perm"]) if "perm" in dbg else np.zeros(0, dtype=np.int64)
            assert perm[lpos:lpos + nl].cpu().numpy().tolist() == o_perm.tolist(), f"call {call} clip {b}"
            if want is not None:
                err = float((c32[lpos:lpos + nl].cpu() - want).abs().max() / want.abs().max())
                assert err < (4e-3 if kmax <= 32 else 2e-3), f"call {call} clip {b}: {err}"
            lpos += nl


@pytest.mark.parametrize("kmax", [32, 40])
def test_lap_kernel_tie_rules_vs_scipy(kmax):
    """tscd_cafm_lap against scipy.optimize.linear_sum_assignment on cost tables FULL of exact ties (small integers,
    duplicated rows/columns), square and rectangular both ways.  kmax = 32 takes the register-resident solver, kmax = 40
    (with frames above 32 rows) the shared-memory one; both must reproduce SciPy's assignment, not just its cost."""
    from scipy.optimize import linear_sum_assignment
    from tscd_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(5)
    shapes = [(1, 1), (1, 7), (7, 1), (5, 5), (30, 30), (32, 32), (30, 17), (17, 30), (32, 3), (3, 32), (29, 31), (31, 29)]
    if kmax > 32:
        shapes += [(40, 40), (33, 20), (20, 33), (36, 40)]
    shapes = shapes * 3
    nf = len(shapes)
    cost = torch.zeros(nf, kmax, kmax)
    for f, (np_, n) in enumerate(shapes):
        kind = f % 3
        if kind == 0:
            c = torch.randint(0, 3, (np_, n), generator=g).float()                 # heavy ties
        elif kind == 1:
            c = torch.randint(0, 4, (np_, 1), generator=g).float() + torch.randint(0, 4, (1, n), generator=g).float()   # rank-1: every assignment optimal
        else:
            c = (torch.rand(np_, n, generator=g) * 8).round() / 8                  # dyadic values, some ties
        cost[f, :np_, :n] = c
    lrow = torch.tensor([0] + list(torch.tensor([n for _, n in shapes]).cumsum(0)), dtype=torch.int32).cuda()
    ref_n = torch.tensor([np_ for np_, _ in shapes], dtype=torch.int32).cuda()
    lap_col = torch.full((nf, kmax), -7, dtype=torch.int32).cuda()
    lap_row = torch.full((nf, kmax), -7, dtype=torch.int32).cuda()
    ops.call("tscd_cafm_lap", L.CafmLapArgs, num_frames=nf, kmax=kmax, lrow_off=lrow, ref_n=ref_n, cost=cost.cuda().contiguous(),
             lap_col=lap_col, lap_row=lap_row)
    torch.cuda.synchronize()
    for f, (np_, n) in enumerate(shapes):
        ri, ci = linear_sum_assignment(cost[f, :np_, :n].double().numpy())
        want_col = [-1] * np_
        want_row = [-1] * n
        for r, c in zip(ri.tolist(), ci.tolist()):
            want_col[r] = c
            want_row[c] = r
        assert lap_col[f, :np_].cpu().tolist() == want_col, f"frame {f} shape {(np_, n)}"
        assert lap_row[f, :n].cpu().tolist() == want_row, f"frame {f} shape {(np_, n)}"


def test_lap_kernel_large_problems_vs_scipy():
    """The shipped OVIS-L limits allow up to 500 proposals per frame (exps/TSCD_OVIS/ovis_tscd_large.py:45,49), i.e. LSAPs of
    up to 500 x 500 (tscd_matching.py:929-935).  tscd_cafm_lap (shared-memory solver, one warp per frame) against
    scipy.optimize.linear_sum_assignment on cosine-like fp32 costs and on a tie-heavy table: identical assignments."""
    import time
    from scipy.optimize import linear_sum_assignment
    from tscd_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(8)
    kmax = 512
    shapes = [(100, 100), (250, 250), (500, 300), (300, 500), (512, 512), (500, 500), (100, 100), (1, 500), (500, 1)]
    nf = len(shapes)
    cost = torch.zeros(nf, kmax, kmax)
    for f, (np_, n) in enumerate(shapes):
        if f == 6:
            c = torch.randint(0, 5, (np_, n), generator=g).float()                      # heavy exact ties
        else:
            a = torch.nn.functional.normalize(torch.randn(np_, 64, generator=g), dim=1)
            b = torch.nn.functional.normalize(torch.randn(n, 64, generator=g), dim=1)
            c = 1.0 - (a @ b.t() + torch.nn.functional.normalize(torch.randn(np_, 32, generator=g), dim=1)
                       @ torch.nn.functional.normalize(torch.randn(n, 32, generator=g), dim=1).t()) / 2
        cost[f, :np_, :n] = c
    lrow = torch.tensor([0] + list(torch.tensor([n for _, n in shapes]).cumsum(0)), dtype=torch.int32).cuda()
    ref_n = torch.tensor([np_ for np_, _ in shapes], dtype=torch.int32).cuda()
    lap_col = torch.full((nf, kmax), -7, dtype=torch.int32).cuda()
    lap_row = torch.full((nf, kmax), -7, dtype=torch.int32).cuda()
    dcost = cost.cuda().contiguous()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ops.call("tscd_cafm_lap", L.CafmLapArgs, num_frames=nf, kmax=kmax, lrow_off=lrow, ref_n=ref_n, cost=dcost,
             lap_col=lap_col, lap_row=lap_row)
    torch.cuda.synchronize()
    print(f"tscd_cafm_lap, {nf} problems up to 512x512: {(time.perf_counter() - t0) * 1e3:.2f} ms")
    for f, (np_, n) in enumerate(shapes):
        ri, ci = linear_sum_assignment(cost[f, :np_, :n].double().numpy())
        want_col = [-1] * np_
        want_row = [-1] * n
        for r, c in zip(ri.tolist(), ci.tolist()):
            want_col[r] = c
            want_row[c] = r
        assert lap_col[f, :np_].cpu().tolist() == want_col, f"frame {f} shape {(np_, n)}"
        assert lap_row[f, :n].cpu().tolist() == want_row, f"frame {f} shape {(np_, n)}"
