#!/usr/bin/env python
"""2+-GPU check of the frame-sharded long-clip mode (torchrun, NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_long_clip.py

Every rank builds the same seeded clip, processes ITS frames (selection on its GPU, one all-gather of the global bank,
CAFM memory handed rank->rank) and rank 0 compares the union of the detections with the single-GPU run of the whole clip."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from tscd_b200 import ops, parallel, selection, stage, weights
    C, D = 25, 256
    hw = [(40, 40), (20, 20), (10, 10)]
    L, G = 4 * world, 12 * world
    Lr, Gr = L // world, G // world
    F = L + G
    an = ops.AnchorSpec(hw)
    A = an.num_anchors
    g = torch.Generator().manual_seed(1234)
    head = torch.cat([torch.rand(F, A, 2, generator=g) * 2 - 0.5, torch.randn(F, A, 2, generator=g) * 0.7 + 1.0,
                      torch.sigmoid(torch.randn(F, A, 1, generator=g) * 2 - 3),
                      torch.sigmoid(torch.randn(F, A, C, generator=g) * 2 - 3)], 2)
    feats = [torch.randn(F, A, D, generator=g).half() for _ in range(3)]
    te = weights.timing_signal_1d(torch.arange(L), 256)
    cfg = stage.StageConfig(num_classes=C, selection=selection.SelectionConfig(mode="A", pre_k=300, top_k=20))
    st = stage.AggregationStage(cfg, weights.random_state_dict(C, D, seed=5), device=dev)
    kmax = cfg.selection.max_keep(A)

    def run_sel(frames):
        h = ops.HeadViews.from_fused(head[frames].cuda(), an, apply_sigmoid=False, apply_decode=True)
        fv = [f[frames].cuda().contiguous() for f in feats]
        views = tuple(ops.view_rowmajor(f, an) for f in fv)
        rows_cap = ((len(frames) * kmax + 127) // 128) * 128 + 128
        return selection.select_and_gather(h, views, torch.float16, D, cfg.selection, bank_dtype=torch.float16,
                                           bank_rows=rows_cap), (h, fv)

    mine = list(range(rank * Lr, (rank + 1) * Lr)) + list(range(L + rank * Gr, L + (rank + 1) * Gr))
    sel, keep_alive = run_sel(mine)
    state = stage.CAFMState(1, kmax, D, dev)
    out = parallel.long_clip_forward(st, sel, Lr, Gr, kmax, te[rank * Lr:(rank + 1) * Lr], state)
    torch.cuda.synchronize()
    res, _ = st.to_lists(out, 1, Lr)
    res = [None if r is None else r.cpu() for r in res]
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    ok = True
    if rank == 0:
        sharded = [r for part in gathered for r in part]
        sel_full, keep2 = run_sel(list(range(F)))
        out_full = st.forward_from_bank(sel_full, 1, F, L, kmax, te, state=stage.CAFMState(1, kmax, D, dev))
        torch.cuda.synchronize()
        full, _ = st.to_lists(out_full, 1, L)
        tot = match = 0
        for a, b in zip(sharded, full):
            if a is None or b is None:
                ok &= (a is None and b is None)
                continue
            b = b.cpu()
            tot += max(len(a), len(b))
            used = set()
            for i in range(len(b)):
                for j in range(len(a)):
                    if j not in used and a[j, 6] == b[i, 6] and torch.allclose(a[j, :6], b[i, :6], rtol=5e-3, atol=0.3):
                        used.add(j); match += 1
                        break
        print(f"long clip sharded over {world} ranks vs single GPU: {match}/{tot} detections identical within tolerance")
        ok &= tot > 0 and match / tot >= 0.99
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
