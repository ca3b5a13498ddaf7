#!/usr/bin/env python
"""WaveletsHFBlock: the selected-anchor evaluation (tscd_edge_patches + per-level tcgen05 GEMMs + tscd_edge_combine)
beside the dense evaluation the reference's head performs (three WaveletsHFBlock(256) modules over the 72/36/18 maps of every
frame; restated here with torch / cuDNN convolutions in fp16 channels_last, the fastest way to run that module on this GPU),
on one 32-frame OVIS clip at 576 x 576, for 30 (mode A) and 500 (mode-B ceiling) proposals per frame.

  python tools/bench_edge.py [--clips 4] [--iters 20]"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HW = [(72, 72), (36, 36), (18, 18)]


def dense_block(x, w1, b1, w3, b3):
    """WaveletsHFBlock.forward (surrounding_extraction.py:257-267) with slicing for the stride-2 Haar steps."""
    a, b = x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2]
    c, d = x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2]
    hf = torch.cat([0.5 * (a + b - c - d), 0.5 * (a - b + c - d), 0.5 * (a - b - c + d)], 1)
    hf = F.relu(F.conv2d(hf, w1, b1))
    C = x.shape[1]
    lh, hl, hh = hf.split([C, C, C], 1)
    out = torch.empty_like(x)
    out[:, :, 0::2, 0::2] = 0.5 * (lh + hl + hh)
    out[:, :, 0::2, 1::2] = 0.5 * (lh - hl - hh)
    out[:, :, 1::2, 0::2] = 0.5 * (-lh + hl - hh)
    out[:, :, 1::2, 1::2] = 0.5 * (-lh - hl + hh)
    return F.relu(F.conv2d(x, w3, b3, padding=1)) * out


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    from tscd_b200 import _lib as L, ops
    dev = "cuda"
    Fn = 32 * args.clips
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(Fn, 256, h, w, generator=g).half().to(dev).contiguous(memory_format=torch.channels_last) for h, w in HW]
    w3 = [(torch.randn(256, 256, 3, 3, generator=g) / 48).half().to(dev).contiguous(memory_format=torch.channels_last) for _ in HW]
    w1 = [(torch.randn(768, 768, 1, 1, generator=g) / 28).half().to(dev) for _ in HW]
    b3 = [torch.randn(256, generator=g).mul(0.1).half().to(dev) for _ in HW]
    b1 = [torch.randn(768, generator=g).mul(0.1).half().to(dev) for _ in HW]
    block = ops.EdgeBlock([w.float() for w in w3], [b.float() for b in b3], [w.float() for w in w1], [b.float() for b in b1], dtype=torch.float16)
    an = ops.AnchorSpec(HW)
    A = an.num_anchors
    t_dense = timed(lambda: [dense_block(feats[l], w1[l], b1[l], w3[l], b3[l]) for l in range(3)], max(3, args.iters // 4))
    flops_dense = 2.0 * Fn * (A * 9 * 256 * 256 + (A // 4) * 768 * 768)
    print(f"dense block (torch / cuDNN fp16 channels_last), {Fn} frames: {t_dense:8.3f} ms   {flops_dense / t_dense / 1e9:7.1f} TFLOP/s")
    for keep in (30, 500):
        idx = torch.stack([torch.randperm(A, generator=g)[:keep] for _ in range(Fn)]).to(torch.int32).to(dev)
        rows_cap = ((Fn * keep + 127) // 128) * 128 + 128
        gathered = dict(bank_edge=torch.empty(rows_cap, 256, dtype=torch.float16, device=dev), max_keep=keep, sel_idx=idx,
                        sel_count=torch.full((Fn,), keep, dtype=torch.int32, device=dev),
                        row_off=(torch.arange(Fn + 1, device=dev) * keep).to(torch.int32))
        view = ops.view_levels(feats)
        t_sel = timed(lambda: ops.edge_rows(block, view, torch.float16, an, gathered, Fn), args.iters)
        # parity against the dense maps on the same rows
        dense = torch.cat([dense_block(feats[l], w1[l], b1[l], w3[l], b3[l]).flatten(2) for l in range(3)], 2).permute(0, 2, 1)
        want = torch.cat([dense[f, idx[f].long()] for f in range(Fn)]).float()
        got = gathered["bank_edge"][:Fn * keep].float()
        err = float((got - want).abs().max() / want.abs().max())
        flops = 2.0 * Fn * keep * (9 * 256 * 256 + 768 * 768)
        print(f"selected anchors, {keep:3d} per frame ({Fn * keep} rows): {t_sel:8.3f} ms   {flops / t_sel / 1e9:7.1f} TFLOP/s   "
              f"{t_dense / t_sel:6.1f}x less time than dense   max err / max |edge| = {err:.1e}")


if __name__ == "__main__":
    main()
