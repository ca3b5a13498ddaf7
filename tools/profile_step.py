#!/usr/bin/env python
"""One profiled replay of a bench configuration (for ncu).  Two warm-up passes run outside the capture range; the third is
bracketed by cudaProfilerStart/Stop, so `ncu --profile-from-start off ...` sees exactly one pass of the stage (eager launches,
classification branch serialised on the launching stream so that ncu's per-launch numbers are for kernels running alone).

  python tools/profile_step.py [--config ovis_a_k30] [--clips 64]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ovis_a_k30", choices=list(bench.CONFIGS))
    ap.add_argument("--clips", type=int, default=0)
    ap.add_argument("--head-layout", default="rows", choices=["rows", "levels"])
    args = ap.parse_args()
    from tscd_b200 import ops, weights
    cfg = bench.CONFIGS[args.config]
    dev = torch.device("cuda", 0)
    B = args.clips or cfg["clips"]
    st, run = bench.make_runner(cfg, dev)
    st.serialize = True
    inp = bench.synth_s1(cfg, B, dev, seed=2024, layout=args.head_layout)
    views = bench.views_of(inp, ops, cfg["C"])
    te = torch.cat([weights.timing_signal_1d(torch.arange(cfg["L"]), 256)] * B, 0).to(dev)
    for _ in range(2):
        out = run(views, B, te)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    out = run(views, B, te)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    assert int(out["status"].item()) == 0
    print("profiled one pass:", args.config, B, "clips")


if __name__ == "__main__":
    main()
