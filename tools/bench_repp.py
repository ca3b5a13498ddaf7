#!/usr/bin/env python
"""REPP post-processing: time of one video through the device-linked REPPB200 beside the CPU restatement of the
reference's REPP class (oracle/repp_oracle.py -- same Python / numpy loops as tools/REPP.py, with the logistic model inlined instead
of one sklearn predict_proba call per detection pair, i.e. FASTER than the reference).  Outputs are compared.

  python tools/bench_repp.py [--frames 120] [--objects 45]"""
import argparse
import copy
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth(T, nobj, seed=5):
    rng = np.random.default_rng(seed)
    ih, iw = 480, 854
    pos = rng.uniform([20, 20], [iw - 150, ih - 150], size=(nobj, 2))
    vel = rng.normal(0, 4, size=(nobj, 2))
    size = rng.uniform(25, 140, size=(nobj, 2))
    cls = rng.integers(0, 30, size=nobj)
    video = {}
    for t in range(T):
        dets = []
        for o in range(nobj):
            for rep in range(int(rng.integers(0, 3))):
                xy = pos[o] + vel[o] * t * 0.3 + rng.normal(0, 2.5, 2)
                wh = size[o] * (1 + rng.normal(0, 0.05, 2))
                x_min, y_min = np.float32(max(0.0, xy[0])), np.float32(max(0.0, xy[1]))
                w_, h_ = np.float32(wh[0]), np.float32(wh[1])
                c = [(x_min + w_ / 2) / max(iw, ih), (y_min + (iw - ih) // 2 + h_ / 2) / max(iw, ih)]
                sc = np.array([rng.uniform(0.05, 0.95), rng.uniform(0.02, 0.9), cls[o] if rep == 0 else int(rng.integers(0, 30))], dtype=np.float32)
                dets.append({"image_id": f"v/{t:06d}", "bbox": [x_min, y_min, w_, h_], "bbox_center": c, "scores": sc})
        video[str(t)] = dets
    return video


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--objects", type=int, default=45)
    args = ap.parse_args()
    import torch
    from oracle import repp_oracle
    from tscd_b200.repp import REPPB200
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "repp.json")))
    cfg = g["cases"][0]["cfg"]
    video = synth(args.frames, args.objects)
    ndet = sum(len(v) for v in video.values())
    post = REPPB200(logreg=g["logreg"], **cfg)
    post(copy.deepcopy(video))                       # warm-up (module load)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = post(copy.deepcopy(video))
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = repp_oracle.REPPOracle(logreg=g["logreg"], **cfg)(copy.deepcopy(video))
    t_cpu = time.perf_counter() - t0
    print(json.dumps({"frames": args.frames, "detections": ndet, "predictions": len(got), "identical": got == want,
                      "repp_b200_s": round(t_dev, 4), "cpu_restatement_s": round(t_cpu, 3), "speedup": round(t_cpu / t_dev, 1)}))


if __name__ == "__main__":
    main()
