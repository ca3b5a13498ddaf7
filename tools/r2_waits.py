#!/usr/bin/env python
"""Debug: where the warp-specialised kernels of one MCA module wait (attn_round2, attn_pv, the fused q|k|v projection).
Needs a library built with TSCD_NVCC_EXTRA=-DTSCD_R2_PROF (python -m tscd_b200.build --force); prints, for CTA (0,0),
the clocks one lane of each role (TMA producer, MMA issuer, softmax / epilogue warp) spent in every barrier wait and the
kernel's total.  Rebuild without the flag afterwards: the counters cost ~20 % of the kernels' time.

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
        63: "kernel total"}
PV_TAGS = {0: "A tma: Q empty", 1: "A tma: kv empty", 10: "A mma: Q full", 11: "A mma: kv full", 12: "A mma: S empty", 20: "A softmax: S full",
           30: "B tma: Q empty", 31: "B tma-k: K empty", 32: "B tma-v: V empty", 40: "B mma-s: Q full", 41: "B mma-s: K full", 42: "B mma-s: S empty",
           45: "B mma-pv: V full", 43: "B mma-pv: P full", 44: "B mma-pv: O empty", 50: "B softmax: S full", 51: "B softmax: P empty",
           52: "B softmax: O full", 62: "pass A total", 63: "kernel total"}


def waits(fn):
    lib = L.lib()
    buf = (ctypes.c_ulonglong * 128)()
    gbuf = (ctypes.c_ulonglong * 8)()
    lib.tscd_debug_r2_waits(buf, 1)
    lib.tscd_debug_gemm_waits(gbuf, 1)
    fn()
    torch.cuda.synchronize()
    lib.tscd_debug_r2_waits(buf, 0)
    lib.tscd_debug_gemm_waits(gbuf, 0)
    waits.gemm = list(gbuf)
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
    g = waits.gemm
    print("\nfused q|k|v projection, CTA 0, both launches of the module summed [clocks]:")
    for i, name in enumerate(["tma: stage empty", "mma: accumulator empty", "mma: stage full", "epilogue warp 2: accumulator full",
                              "total tma warp", "total mma warp", "total epilogue warp 2"]):
        print(f"{name:36s} {g[i]:12d}")
    print(f"\n{'attn_pv wait site':28s} {'need_reg=0':>12s} {'need_reg=1':>12s}")
    for t, name in PV_TAGS.items():
        print(f"{name:28s} {cls[64 + t]:12d} {both[64 + t]:12d}")


if __name__ == "__main__":
    main()
