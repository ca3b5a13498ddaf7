#!/usr/bin/env python
"""Micro-benchmark of one MCA module (prep -> pv -> round2 + its GEMMs) at the bench shape, each C-ABI entry timed
alone with CUDA events on one stream (no side-stream overlap).  Synthetic bank: B clips x 32 frames x 30 proposals.

  python tools/bench_mca.py [--clips 64] [--reps 5] [--need-reg 1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscd_b200 import _lib as L, aggregate, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--need-reg", type=int, default=1)
    ap.add_argument("--k", type=int, default=30)
    ap.add_argument("--bf16", action="store_true")
    args = ap.parse_args()
    B, F, Lf, K, D = args.clips, 32, 8, args.k, 256
    dt = torch.bfloat16 if args.bf16 else torch.float16
    g = torch.Generator(device="cuda").manual_seed(1)
    N = B * F * K
    row_cap = (N + 127) // 128 * 128 + 128
    loc_cap = (B * Lf * K + 127) // 128 * 128
    cnt = torch.full((B * F,), K, dtype=torch.int32, device="cuda")
    lay = aggregate.make_layout(cnt, B, F, Lf, row_cap, loc_cap, (F * K + 127) // 128 * 128, dt)
    bank_c = torch.randn(row_cap, D, generator=g, device="cuda").to(dt)
    bank_r = torch.randn(row_cap, D, generator=g, device="cuda").to(dt)
    score = torch.rand(row_cap, generator=g, device="cuda")
    sd = weights.random_state_dict(25, D, seed=3)
    w = aggregate.MCAWeights(sd, "agg_iou.", dt)
    n_dev, nl_dev = lay.row_off[-1:], lay.lrow_off[-1:]
    for _ in range(2):
        aggregate.mca_forward(lay, w, bank_c, bank_r, score, n_dev, nl_dev, need_reg=bool(args.need_reg))
    torch.cuda.synchronize()
    L.profile = {"names": None, "events": {}}
    for _ in range(args.reps):
        aggregate.mca_forward(lay, w, bank_c, bank_r, score, n_dev, nl_dev, need_reg=bool(args.need_reg))
    torch.cuda.synchronize()
    tot = 0.0
    for k, v in L.profile["events"].items():
        ms = [s.elapsed_time(e) for s, e in v]
        per = len(ms) // args.reps
        tot += sum(ms) / args.reps
        print(f"{k:20s} {per} launches/module  {1e3 * sum(ms) / len(ms):8.1f} us/launch  {1e3 * sum(ms) / args.reps:8.1f} us/module")
    print(f"module total {1e3 * tot:.1f} us  ({B} clips, {N} bank rows)")


if __name__ == "__main__":
    main()
