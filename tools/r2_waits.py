#!/usr/bin/env python
"""Debug: where the round-2 attention kernel waits.  Needs a library built with TSCD_NVCC_EXTRA=-DTSCD_R2_PROF
(python -m tscd_b200.build --force); prints, for CTA (0,0), the clocks one lane of each role spent in every barrier wait.

  TSCD_NVCC_EXTRA=-DTSCD_R2_PROF python -m tscd_b200.build --force && python tools/r2_waits.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscd_b200 import _lib as L, aggregate, weights  # noqa: E402

TAGS = {0: "tma-s: slot empty", 1: "tma-x: slot empty", 10: "mma-s: item full", 12: "mma-s: S empty", 15: "mma-x: item full",
        11: "mma-x: R empty", 13: "mma-x: W full", 14: "mma: resident full", 20: "softmax: R full", 21: "softmax: S full", 22: "softmax: W empty", 23: "softmax: U full",
        31: "kernel total"}


def waits(fn):
    lib = L.lib()
    buf = (ctypes.c_ulonglong * 32)()
    lib.tscd_debug_r2_waits(buf, 1)
    fn()
    torch.cuda.synchronize()
    lib.tscd_debug_r2_waits(buf, 0)
    return list(buf)


def main():
    B, F, Lf, K, D = 64, 32, 8, 30, 256
    dt = torch.float16
    g = torch.Generator(device="cuda").manual_seed(1)
    N = B * F * K
    row_cap = (N + 127) // 128 * 128 + 128
    loc_cap = (B * Lf * K + 127) // 128 * 128
    cnt = torch.full((B * F,), K, dtype=torch.int32, device="cuda")
    lay = aggregate.make_layout(cnt, B, F, Lf, row_cap, loc_cap, (F * K + 127) // 128 * 128, dt)
    bank_c = torch.randn(row_cap, D, generator=g, device="cuda").to(dt)
    bank_r = torch.randn(row_cap, D, generator=g, device="cuda").to(dt)
    score = torch.rand(row_cap, generator=g, device="cuda")
    w = aggregate.MCAWeights(weights.random_state_dict(25, D, seed=3), "agg_iou.", dt)
    n_dev, nl_dev = lay.row_off[-1:], lay.lrow_off[-1:]
    res = {}
    for need_reg in (False, True):
        f = lambda: aggregate.mca_forward(lay, w, bank_c, bank_r, score, n_dev, nl_dev, need_reg=need_reg)
        f(); f()
        res[need_reg] = waits(f)
    cls, both = res[False], res[True]
    print(f"{'wait site':28s} {'cls launch':>12s} {'obj launch (w_in)':>18s}   [clocks, CTA (0,0)]")
    for t, name in TAGS.items():
        print(f"{name:28s} {cls[t]:12d} {both[t] - cls[t]:18d}")


if __name__ == "__main__":
    main()
