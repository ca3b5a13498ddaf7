"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: clip sharding, the ragged all-gather of the global
bank for the frame-sharded long clip, and the CAFM state hand-over between ranks."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_clips_contiguous_and_complete():
    from tscd_b200.parallel import shard_clips
    for n in (1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_clips(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _make_rank_sel(rank, Lr, Gr, kmax, D=8):
    g = torch.Generator().manual_seed(100 + rank)
    counts = torch.randint(1, kmax + 1, (Lr + Gr,), generator=g).tolist()
    n = sum(counts)
    cap = ((Lr + Gr) * kmax + 127) // 128 * 128 + 128
    sel = {}
    for k in ("bank_cls", "bank_reg", "bank_edge"):
        t = torch.zeros(cap, D)
        t[:n] = torch.randn(n, D, generator=g) + 10 * rank
        sel[k] = t
    sc = torch.zeros(cap)
    sc[:n] = torch.rand(n, generator=g)
    sel["bank_score"] = sc
    sel["sel_count"] = torch.tensor(counts, dtype=torch.int32)
    ro = torch.zeros(Lr + Gr + 1, dtype=torch.int32)
    ro[1:] = torch.cumsum(sel["sel_count"], 0)
    sel["row_off"] = ro
    sel["sel_rows"] = torch.zeros(Lr + Gr, kmax, 9)
    return sel, counts


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tscd_b200 import parallel
        Lr, Gr, kmax = 2, 3, 5
        sel, counts = _make_rank_sel(rank, Lr, Gr, kmax)
        virt, F_virt = parallel.exchange_global_bank(sel, Lr, Gr, kmax)
        # expected virtual clip, rebuilt from every rank's deterministic data
        all_sel = [_make_rank_sel(r, Lr, Gr, kmax) for r in range(world)]
        exp_counts = counts[:Lr] + [c for (_, cs) in all_sel for c in cs[Lr:]]
        assert F_virt == Lr + world * Gr
        assert virt["sel_count"].tolist() == exp_counts
        assert virt["row_off"].tolist() == [0] + torch.cumsum(torch.tensor(exp_counts), 0).tolist()
        n_loc = sum(counts[:Lr])
        for k in ("bank_cls", "bank_reg", "bank_edge", "bank_score"):
            parts = [sel[k][:n_loc]]
            for (s_r, cs) in all_sel:
                lo = sum(cs[:Lr])
                parts.append(s_r[k][lo:lo + sum(cs[Lr:])])
            exp = torch.cat(parts)
            assert torch.equal(virt[k][:exp.shape[0]], exp), k
        # CAFM state hand-over rank 0 -> rank 1
        class St:
            pass
        st = St()
        for i, f in enumerate(parallel._STATE_FIELDS):
            setattr(st, f, torch.full((3, 4), float(i + 1) if rank == 0 else 0.0) if f != "n" else
                    torch.tensor([7 if rank == 0 else 0], dtype=torch.int32))
        if rank == 0:
            parallel.send_state(st, 1)
        else:
            parallel.recv_state(st, 0)
            assert int(st.n.item()) == 7 and float(st.time[0, 0]) == float(len(parallel._STATE_FIELDS))
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_long_clip_exchange_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
