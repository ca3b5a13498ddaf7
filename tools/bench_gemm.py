#!/usr/bin/env python
"""Micro-benchmark of tscd_linear on the stage's GEMM shapes (CUDA events, L2-cold by rotating buffers)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscd_b200 import ops  # noqa: E402

SHAPES = [(61568, 768, 256, "qkv (bank rows)"), (15360, 512, 512, "mca.linear"), (15360, 1024, 768, "out 768->1024"),
          (15360, 1024, 256, "fc_reg_matcher"), (15360, 1024, 1024, "ta_q"), (15360, 2048, 1024, "ta_kv"),
          (15360, 256, 256, "cafm k/v"), (15360, 25, 1024, "cls_pred")]


def main():
    shapes = SHAPES[:1] if os.environ.get('ONLY_FIRST') else SHAPES
    for M, N, K, name in shapes:
        nbuf = 6
        xs = [torch.randn(M, K, device="cuda").half() for _ in range(nbuf)]
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        outs = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(nbuf)]
        for i in range(3):
            ops.linear(xs[i % nbuf], w, out16=outs[i % nbuf], want16=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if os.environ.get('ONLY_FIRST') else 30
        e0.record()
        for i in range(reps):
            ops.linear(xs[i % nbuf], w, out16=outs[i % nbuf], want16=False)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        flops = 2.0 * M * N * K
        byts = (M * K + N * K + M * N) * 2
        ref = xs[0].float() @ w.float().t()
        err = float((outs[0].float() - ref).abs().max() / ref.abs().max())
        print(f"{name:18s} M={M} N={N} K={K}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {byts / us / 1e3:7.1f} GB/s  err {err:.1e}")


if __name__ == "__main__":
    main()
