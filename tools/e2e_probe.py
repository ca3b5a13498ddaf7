#!/usr/bin/env python
"""Where does a forward_host() call spend its time?  Phases of the host-buffer path on one GPU:

  python tools/e2e_probe.py [--clips 32] [--chunk 8]

Prints, for the bench's headline workload with pinned HOST inputs: wall time per call, GPU time of the captured graph (CUDA
events around the replay), the host-side unpack, and K1 / K3 alone on the host-resident rows vs on device copies."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=32)
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--config", default="ovis_a_k30")
    ap.add_argument("--phases-only", action="store_true")
    ap.add_argument("--pipeline-only", action="store_true")
    ap.add_argument("--pre-k", type=int, default=0, help="override the configuration's pre_k (what-if timing only)")
    args = ap.parse_args()
    from tscd_b200 import ops, selection, weights
    cfg = dict(bench.CONFIGS[args.config])
    if args.pre_k:
        cfg["pre_k"] = args.pre_k
    dev = torch.device("cuda", 0)
    F, Lf, C = cfg["F"], cfg["L"], cfg["C"]
    Be = args.clips
    st, run = bench.make_runner(cfg, dev)
    dev_src = bench.synth_s1(cfg, Be, dev, seed=99, layout="rows")
    host = {k: [(torch.empty(t.shape, dtype=t.dtype, pin_memory=True, memory_format=torch.channels_last) if t.dim() == 4 else
                 torch.empty(t.shape, dtype=t.dtype, pin_memory=True)).copy_(t) for t in v] for k, v in dev_src.items()}
    te = torch.cat([weights.timing_signal_1d(torch.arange(Lf), 256)] * Be, 0).pin_memory()

    def timed(fn, n=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / n

    for graph, lanes, chunk in [] if (args.phases_only or args.pipeline_only) else [(True, 2, 8), (True, 1, 8), (True, 3, 8), (True, 4, 8), (True, 2, 4), (True, 4, 4), (True, 2, 16), (True, 1, 32),
                                (False, 2, 8), (False, 2, 16)]:
        ms = timed(lambda: st.forward_host(host, bench.HW, te, Be, F, Lf, chunk_clips=chunk, graph=graph, lanes=lanes))
        plan = list(st._host_plans.values())[-1]
        gpu = ""
        if plan["graph"] is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            plan["graph"].replay()
            e1.record()
            torch.cuda.synchronize()
            gpu = f" (GPU {e0.elapsed_time(e1):.2f} ms)"
        print(f"forward_host graph={graph} lanes={lanes} chunk={chunk}: {ms:.2f} ms/call{gpu} -> {Be * F / ms * 1e3:.0f} clip-frames/s")
    # pipelined throughput: `depth` calls in flight on separate slots
    for depth, lanes, chunk in [] if args.phases_only else [(2, 4, 8), (3, 4, 8), (3, 6, 4), (4, 4, 8), (3, 4, 4), (4, 6, 4), (3, 8, 4), (3, 8, 2)]:
        def submit(i):
            return st.forward_host_submit(host, bench.HW, te, Be, F, Lf, chunk_clips=chunk, lanes=lanes, slot=i % depth)
        for i in range(2 * depth):
            st.forward_host_collect(submit(i))
        n, infl = 12 * depth, []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            infl.append(submit(i))
            if len(infl) == depth:
                st.forward_host_collect(infl.pop(0))
        while infl:
            st.forward_host_collect(infl.pop(0))
        ms = 1e3 * (time.perf_counter() - t0) / n
        print(f"pipelined depth={depth} lanes={lanes} chunk={chunk}: {ms:.2f} ms/call -> {Be * F / ms * 1e3:.0f} clip-frames/s")
        st._host_plans.clear()
    if args.pipeline_only:
        return
    # phases of the graph path
    st.forward_host(host, bench.HW, te, Be, F, Lf, chunk_clips=args.chunk, graph=True)
    plan = list(st._host_plans.values())[-1]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    plan["graph"].replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph replay GPU time: {e0.elapsed_time(e1):.2f} ms")
    t0 = time.perf_counter()
    for nc, pk in plan["pend"]:
        st._unpack_host(pk, nc * Lf)
    print(f"host unpack: {1e3 * (time.perf_counter() - t0):.2f} ms")
    # K1 / K3 alone
    an = ops.AnchorSpec(bench.HW)
    objp_dev = host["objp"][0].to(dev)
    sel_cfg = bench.selection_of(cfg, selection)
    feats_h = tuple(ops.view_levels(host[k]) for k in ("f_cls", "f_reg", "f_edge"))
    feats_d = tuple(ops.view_levels(dev_src[k]) for k in ("f_cls", "f_reg", "f_edge"))
    for name, rows, feats in (("host rows + host features", host["rows"][0], feats_h), ("device rows + host features", dev_src["rows"][0], feats_h),
                              ("device rows + device features", dev_src["rows"][0], feats_d)):
        head = ops.HeadViews.from_rows(rows, objp_dev, an, C)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for _ in range(2):
            ev[0].record()
            cand = ops.select(head, "A", pre_k=cfg["pre_k"])
            ev[1].record()
            keep, kc, _ = ops.nms(cand["box"], cand["score"], cand["cls"], cand["count"], 0.75, max_keep=cfg["top_k"])
            ev[2].record()
            ops.gather(head, feats, torch.float16, 256, cand, keep, kc, max_keep=cfg["top_k"], bank_dtype=torch.float16)
            ev[3].record()
            torch.cuda.synchronize()
        print(f"{name}: select {ev[0].elapsed_time(ev[1]):.3f} ms, nms {ev[1].elapsed_time(ev[2]):.3f} ms, gather {ev[2].elapsed_time(ev[3]):.3f} ms "
              f"({Be * F} frames)")


if __name__ == "__main__":
    main()
