#!/usr/bin/env python
"""One profiled step of the bench workload (for ncu).  Two warm-up steps run outside the capture range; the
third step is bracketed by cudaProfilerStart/Stop, so `ncu --profile-from-start off ...` sees exactly one step.

  python tools/profile_step.py [--clips 16]
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
    ap.add_argument("--clips", type=int, default=16)
    args = ap.parse_args()
    from tscd_b200 import ops, selection, stage, weights
    dev = torch.device("cuda", 0)
    B = args.clips
    cfg = stage.StageConfig(num_classes=bench.C, selection=selection.SelectionConfig(mode="A", pre_k=bench.PRE_K, top_k=bench.TOP_K))
    st = stage.AggregationStage(cfg, weights.random_state_dict(bench.C, bench.D, seed=2024), device=dev)
    inp = bench.synth_s1(B, dev, seed=2024)
    head, feats = bench.views_of(inp, ops)
    te = torch.cat([weights.timing_signal_1d(torch.arange(bench.LF), 256)] * B, 0).to(dev)
    for _ in range(2):
        out = st.forward(head, feats, torch.float16, te, B, bench.F, bench.LF)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    out = st.forward(head, feats, torch.float16, te, B, bench.F, bench.LF)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    res, _ = st.to_lists(out, B, bench.LF)
    print("profiled one step:", B, "clips,", sum(0 if r is None else len(r) for r in res), "detections")


if __name__ == "__main__":
    main()
