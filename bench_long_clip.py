"""bench.py --mode long-clip (BASELINE.json configs[3]): ONE 256-frame OVIS clip sharded by frame over the ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P bench.py --gpus 8 --mode long-clip

L = 64 local + G = 192 global frames (the OVIS-L val ratio 8 : 24), 25 classes, 576x576, top-750 -> NMS 0.75 -> 30 proposals per
frame.  Every rank owns L/W consecutive local and G/W global frames: K1-K3 on its own frames, ONE all-gather of the packed
global-frame bank rows (tscd_bank_pack -> ncclAllGather -> tscd_bank_unpack), attention of its local rows against
[own frame | all 192 x 30 global rows], CAFM chain pipelined rank -> rank (one P2P message per hop), tail + final NMS.
A step = `--replays` clips processed back to back.  Reports clip-frames/s, the per-phase device times of rank 0 (K1-K3,
exchange, rest) and the time rank W-1 waits for the CAFM memory (the pipeline bubble), and -- once, outside the timed region --
compares the sharded detections with the same clip run in one piece on rank 0's GPU."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_TOTAL, G_TOTAL, C, D = 64, 192, 25, 256


def _frame_inputs(bench, frames, dev, clip_seed):
    """Seam S1 tensors of the given frames of the clip (deterministic per frame, so any rank can rebuild any frame)."""
    cfg = dict(bench.CONFIGS["ovis_a_k30"], F=1)
    parts = [bench.synth_s1(cfg, 1, dev, seed=clip_seed * 100003 + f) for f in frames]
    return {k: [torch.cat([p[k][l] for p in parts], 0).contiguous(memory_format=torch.channels_last) for l in range(3)] for k in parts[0]}


def main(args, rank, world, local):
    import bench
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: tscd_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from tscd_b200 import _lib as Lb, ops, parallel, selection, stage, weights
    sel_cfg = selection.SelectionConfig(mode="A", pre_k=750, top_k=30, nms_thresh=0.75)
    st = stage.AggregationStage(stage.StageConfig(num_classes=C, selection=sel_cfg), weights.random_state_dict(C, D, seed=2024), device=dev)
    kmax = 30
    Lr, Gr = L_TOTAL // world, G_TOTAL // world
    loc, glob = parallel.frame_plan(L_TOTAL, G_TOTAL, rank, world)
    n_clips = 2                                               # rotating input clips
    inputs = [_frame_inputs(bench, loc + glob, dev, clip_seed=7 + c) for c in range(n_clips)]
    views = [bench.views_of(i, ops) for i in inputs]
    te_all = weights.timing_signal_1d(torch.arange(L_TOTAL), 256).to(dev)
    te = te_all[rank * Lr:(rank + 1) * Lr].contiguous()
    state = stage.CAFMState(1, kmax, D, dev)
    rows_cap = (((Lr + Gr) * kmax + 127) // 128) * 128 + 128
    ev = lambda: torch.cuda.Event(enable_timing=True)         # noqa: E731
    phases = []

    def one_clip(i, timed=False):
        head, feats = views[i % n_clips]
        e = [ev() for _ in range(4)] if timed else None
        if timed:
            e[0].record()
        sel = selection.select_and_gather(head, feats, torch.float16, D, sel_cfg, bank_dtype=torch.float16, bank_rows=rows_cap)
        if timed:
            e[1].record()
        if world > 1:
            virt, F_virt = parallel.exchange_global_bank(sel, Lr, Gr, kmax, torch.float16)
        else:
            virt, F_virt = sel, Lr + Gr
        if timed:
            e[2].record()
        resume = torch.tensor([1 if rank > 0 else 0], dtype=torch.int32, device=dev)
        out = st.forward_from_bank(virt, 1, F_virt, Lr, kmax, te, state=state, resume=resume,
                                   before_cafm=(lambda s_: parallel.recv_state(s_, rank - 1)) if rank > 0 else None,
                                   after_cafm=(lambda s_: parallel.send_state(s_, rank + 1)) if rank < world - 1 else None)
        if timed:
            e[3].record()
            phases.append(e)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    R = args.replays or 16
    for i in range(max(3, args.warmup)):
        out = one_clip(i)
    barrier()
    st.to_lists(out, 1, Lr)
    Lb.launch_count = 0
    out = one_clip(0)
    launches = Lb.launch_count
    clocks = bench.ClockSampler(local)
    clocks.start()
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for s_ in range(args.steps):
        for r in range(R):
            out = one_clip(r, timed=(s_ == 0))
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ph = [[e[k].elapsed_time(e[k + 1]) for k in range(3)] for e in phases]
    ph_mean = [sum(p[k] for p in ph) / len(ph) for k in range(3)]
    allph = [None] * world
    if world > 1:
        dist.all_gather_object(allph, ph_mean)
    else:
        allph = [ph_mean]
    # ---- parity: the sharded clip vs the same clip in one piece on rank 0 ----
    res, res_ori = st.to_lists(one_clip(0), 1, Lr)
    res = [None if r is None else r.cpu() for r in res]
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, res)
    else:
        gathered = [res]
    parity = None
    if rank == 0:
        del inputs, views
        torch.cuda.empty_cache()
        full_in = _frame_inputs(bench, list(range(L_TOTAL + G_TOTAL)), dev, clip_seed=7)
        head, feats = bench.views_of(full_in, ops)
        out_full = st.forward(head, feats, torch.float16, te_all, 1, L_TOTAL + G_TOTAL, L_TOTAL)
        torch.cuda.synchronize()
        full, _ = st.to_lists(out_full, 1, L_TOTAL)
        sharded = [r for part in gathered for r in part]
        tot = match = 0
        for a, b in zip(sharded, full):
            if a is None or b is None:
                continue
            b = b.cpu()
            tot += max(len(a), len(b))
            used = set()
            for i in range(len(b)):
                same = torch.where(a[:, 6] == b[i, 6])[0].tolist()
                for j in same:
                    if j not in used and float((a[j, :4] - b[i, :4]).abs().max()) <= 0.3 and torch.allclose(a[j, 4:6], b[i, 4:6], rtol=1e-2, atol=1e-5):
                        used.add(j); match += 1
                        break
        parity = {"matched": match, "total": tot, "frac": match / max(1, tot), "tolerance": "boxes 0.3 px, scores 1e-2 relative"}
        frames = L_TOTAL + G_TOTAL
        line = {"metric": bench.METRIC, "value": frames * R * args.steps / (ms / 1e3), "unit": "clip-frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"name": "long_clip", "mode": "long-clip",
                           "workload": f"ONE {frames}-frame OVIS clip ({L_TOTAL} local + {G_TOTAL} global), 25cls @576x576, top-750 -> NMS0.75 -> 30/frame, "
                                       f"sharded by frame over {world} rank(s); {R} clips per step back to back",
                           "parallelism": f"frame-sharded x{world}: one ncclAllGather of the packed global bank ({Lb.lib().tscd_bank_pack_bytes(Gr, kmax)} B per rank) "
                                          "+ CAFM memory handed rank->rank (one P2P message per hop)"},
                "clocks": clk, "gpu_launches": launches * R * args.steps,
                "long_clip": {"ms_per_clip": ms / args.steps / R,
                              "phase_ms_per_rank": [{"rank": r, "k1_k3": round(p[0], 4), "pack_allgather_unpack": round(p[1], 4),
                                                     "attention_cafm_tail_incl_wait_for_previous_rank": round(p[2], 4)} for r, p in enumerate(allph)],
                              "parity_vs_single_gpu": parity}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
