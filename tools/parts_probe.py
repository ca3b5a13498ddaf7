#!/usr/bin/env python
"""Does running independent sub-batches of clips on concurrent streams (AggregationStage.forward_concurrent) beat one launch
sequence?  The latency-bound kernels of one sub-batch (CAFM chain, LSAP, NMS) can overlap the tensor-core kernels of another.

  python tools/parts_probe.py [--clips 148] [--parts 1,2,4]
Prints clip-frames/s of a CUDA-graph replay for every split of the same inputs."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=148)
    ap.add_argument("--parts", default="1,2,4")
    ap.add_argument("--config", default="ovis_a_k30")
    args = ap.parse_args()
    from tscd_b200 import ops, weights
    cfg = bench.CONFIGS[args.config]
    dev = torch.device("cuda", 0)
    F, Lf, C = cfg["F"], cfg["L"], cfg["C"]
    B = args.clips
    st, run = bench.make_runner(cfg, dev)
    inp = bench.synth_s1(cfg, B, dev, seed=2024, layout="rows")
    te = torch.cat([weights.timing_signal_1d(torch.arange(Lf), 256)] * B, 0).to(dev)
    an = ops.AnchorSpec(bench.HW)
    for nparts in [int(x) for x in args.parts.split(",")]:
        bounds = [(B * i) // nparts for i in range(nparts + 1)]
        parts = []
        for i in range(nparts):
            c0, c1 = bounds[i], bounds[i + 1]
            f0, f1 = c0 * F, c1 * F
            head = ops.HeadViews.from_rows(inp["rows"][0][f0:f1], inp["objp"][0][f0:f1], an, C)
            feats = tuple(ops.view_levels([t[f0:f1] for t in inp[k]]) for k in ("f_cls", "f_reg", "f_edge"))
            parts.append((head, feats, te[c0 * Lf:c1 * Lf], c1 - c0))

        def fn():
            if nparts == 1:
                h, f, t, nb = parts[0]
                return [st.forward(h, f, torch.float16, t, nb, F, Lf)]
            return st.forward_concurrent(parts, torch.float16, F, Lf)

        g, outs = st.capture_fn(fn)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        assert all(int(o["status"].item()) == 0 for o in outs)
        print(f"clips {B} parts {nparts}: {ms * 1e3:.0f} us/replay -> {B * F / ms * 1e3:.0f} clip-frames/s")
        del g, outs


if __name__ == "__main__":
    main()
