#!/usr/bin/env python
"""Generate the host-side goldens (tests/golden/clips.json, repp.json, edge.npz) by RUNNING the reference's own code.

Build container only (needs /root/reference):  python tools/make_goldens_host.py
No reference source is copied: the functions are imported (or, for the clip loop that tscd_demo.py has inline in
`imageflow_demo`, the lines are read from the checkout and executed in place with an index list standing in for the decoded frames)
and their inputs / outputs recorded.

clips.json   OVIS.photo_to_sequence (yolox/data/datasets/vid.py:601-683), VIDDataset.photo_to_sequence (:133-237) and the clip
             loop of tools/tscd_demo.py:209-253 under seeded `random`, on synthetic videos of assorted lengths.
repp.json    tools/REPP.py REPP.__call__ on seeded synthetic detections of a 14-frame video (moving objects, detections typed like
             Predictor.to_repp_heavy's: numpy float32 scalars), for the shipped configuration (tools/yolo_repp_cfg.json: logreg /
             dot, re-coordination) and variants (def distance, add_unmatched, no re-coordination); plus the coefficients of the
             linking model (tools/matching_model_logreg.pckl).
edge.npz     yolox/models/surrounding_extraction.py WaveletsHFBlock.forward on seeded inputs and weights (fp32, CPU)."""
import copy
import json
import os
import pickle
import random
import sys
import tempfile
import textwrap
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

ref_shim.install()


def gen_clips():
    from yolox.data.datasets import vid as V
    cases = []
    lengths = [5, 31, 32, 33, 40, 64, 77, 100]
    # OVIS.photo_to_sequence reads a COCO-style json: videos + images(sid, file_name)
    for mode, lf, gf in (("random", 8, 24), ("random", 0, 32), ("uniform", 0, 16), ("gl", 4, 12), ("random", 1, 31)):
        anno = {"videos": [{"id": i} for i in range(len(lengths))], "images": []}
        for sid, n in enumerate(lengths):
            for k in range(n):
                anno["images"].append({"sid": sid, "file_name": f"v{sid}/img_{k:07d}.jpg"})
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(anno, f)
        fake = types.SimpleNamespace(coco_anno_path=f.name, mode=mode, training=False, val=True)
        random.seed(1234)
        res = V.OVIS.photo_to_sequence(fake, lf, gf)
        os.unlink(f.name)
        cases.append(dict(fn="ovis", mode=mode, lframe=lf, gframe=gf, seed=1234, lengths=lengths, clips=res))
    for mode, lf, gf, formal in (("random", 1, 31, False), ("random", 4, 12, True), ("random", 0, 32, True), ("uniform", 0, 16, False)):
        videos = [[f"v{sid}/{k:06d}.JPEG" for k in range(n)] for sid, n in enumerate(lengths)]
        fake = types.SimpleNamespace(mode=mode, training=False, val=True, formal=formal, local_stride=1, traj_linking=False, tnum=-1)
        random.seed(99)
        res = V.VIDDataset.photo_to_sequence(fake, copy.deepcopy(videos), lf, gf)
        cases.append(dict(fn="vid", mode=mode, lframe=lf, gframe=gf, formal=formal, seed=99, lengths=lengths, clips=res))
    # tools/tscd_demo.py: the clip loop is inline in imageflow_demo -- execute exactly those lines of the checkout
    src = open(os.path.join(REF, "tools", "tscd_demo.py")).read().split("\n")
    first = next(i for i, l in enumerate(src) if l.strip() == "res = []" and "path_sequence = []" in src[i + 1])
    last = next(i for i in range(first, len(src)) if src[i].strip().startswith("outputs, adj_lists, fc_outputs, names"))
    body = textwrap.dedent("\n".join(src[first:last]))
    for n, lf, gf in ((100, 8, 24), (37, 8, 24), (20, 8, 24), (64, 0, 32), (50, 4, 0), (9, 4, 12)):
        env = dict(frames=list(range(n)), random=random, lframe=lf, gframe=gf, traj_linking=False)
        random.seed(7)
        exec(body, env)
        cases.append(dict(fn="demo", frame_len=n, lframe=lf, gframe=gf, seed=7, clips=env["res"], path_sequence=env["path_sequence"]))
    json.dump(cases, open(os.path.join(OUT, "clips.json"), "w"))


def synth_video(rng, frames=14, objects=9, classes=30):
    """Detections typed like Predictor.to_repp_heavy (tools/val_to_imdb.py:193-218): numpy float32 scalars."""
    ih, iw = 360, 640
    wd, hd = max(0, (ih - iw) // 2), max(0, (iw - ih) // 2)
    pos = rng.uniform([40, 40], [iw - 120, ih - 120], size=(objects, 2))
    vel = rng.normal(0, 6, size=(objects, 2))
    size = rng.uniform(30, 110, size=(objects, 2))
    cls = rng.integers(0, 6, size=objects)
    video = {}
    for t in range(frames):
        dets = []
        for o in range(objects):
            if rng.random() < 0.12:
                continue                                         # missed detection
            for rep in range(1 if rng.random() < 0.7 else 2):    # occasional duplicate (another class hypothesis)
                xy = pos[o] + vel[o] * t + rng.normal(0, 2.0, 2)
                wh = size[o] * (1 + rng.normal(0, 0.04, 2))
                out = np.array([xy[0], xy[1], xy[0] + wh[0], xy[1] + wh[1], rng.uniform(0.2, 0.95), rng.uniform(0.05, 0.9),
                                cls[o] if rep == 0 else (cls[o] + 1) % classes], dtype=np.float32)
                x_min, y_min, x_max, y_max = out[:4]
                y_min, x_min = max(0, y_min), max(0, x_min)
                y_max, x_max = min(ih, y_max), min(iw, x_max)
                width, height = x_max - x_min, y_max - y_min
                if width <= 0 or height <= 0:
                    continue
                center = [(x_min + wd + width / 2) / max(iw, ih), (y_min + hd + height / 2) / max(iw, ih)]
                dets.append({"image_id": f"vid0/{t:06d}", "bbox": [x_min, y_min, width, height], "bbox_center": center, "scores": out[4:7]})
        for _ in range(rng.integers(0, 4)):                       # clutter
            out = np.array([rng.uniform(0, iw - 60), rng.uniform(0, ih - 60), 0, 0, rng.uniform(0.01, 0.3), rng.uniform(0.01, 0.3),
                            rng.integers(0, classes)], dtype=np.float32)
            w_, h_ = np.float32(rng.uniform(10, 60)), np.float32(rng.uniform(10, 60))
            x_min, y_min = out[0], out[1]
            center = [(x_min + wd + w_ / 2) / max(iw, ih), (y_min + hd + h_ / 2) / max(iw, ih)]
            dets.append({"image_id": f"vid0/{t:06d}", "bbox": [x_min, y_min, w_, h_], "bbox_center": center, "scores": out[4:7]})
        video[str(t)] = dets
    return video


def jsonable_video(video):
    # a clamped coordinate is the Python int 0 (max(0, y_min)): keep it an int -- numpy's dtype propagation depends on it
    return {k: [{"image_id": p["image_id"], "bbox": [v if isinstance(v, int) else float(v) for v in p["bbox"]], "bbox_center": [float(v) for v in p["bbox_center"]],
                 "scores": [float(v) for v in p["scores"]]} for p in v] for k, v in video.items()}


def gen_repp():
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, "gaussian"):                    # removed in SciPy 1.13; the reference still calls signal.gaussian
        scipy.signal.gaussian = scipy.signal.windows.gaussian
    sys.path.insert(0, os.path.join(REF, "tools"))
    import REPP as R
    warnings.filterwarnings("ignore")
    model, feats = pickle.load(open(os.path.join(REF, "tools", "matching_model_logreg.pckl"), "rb"))
    logreg = dict(features=list(feats), coef=np.asarray(model.coef_).reshape(-1).tolist(), intercept=float(model.intercept_[0]))
    base = json.load(open(os.path.join(REF, "tools", "yolo_repp_cfg.json")))
    base["weight_path"] = os.path.join(REF, "tools", "matching_model_logreg.pckl")
    variants = [dict(), dict(distance_func="def"), dict(add_unmatched=True), dict(recoordinate=False, clf_mode="max"),
                dict(clf_mode="dot_plus", clf_thr=0.5, min_tubelet_score=0.02)]
    cases = []
    for vi, var in enumerate(variants):
        cfg = dict(base, **var)
        video = synth_video(np.random.default_rng(100 + vi))
        inp = jsonable_video(video)
        out = R.REPP(**cfg)(copy.deepcopy(video))
        cfg_j = {k: v for k, v in cfg.items() if k != "weight_path"}
        cases.append(dict(cfg=cfg_j, video=inp, out=[{k: (v if not isinstance(v, (np.floating, np.integer)) else v.item()) for k, v in o.items()} for o in out]))
        print("repp variant", vi, var, "->", len(out), "predictions,", len({o["track_id"] for o in out}), "tracks")
    json.dump(dict(logreg=logreg, cases=cases), open(os.path.join(OUT, "repp.json"), "w"))


def gen_edge():
    """WaveletsHFBlock (yolox/models/surrounding_extraction.py:215-267): seeded inputs / weights -> the module's output (fp32, CPU)."""
    import torch
    from yolox.models.surrounding_extraction import WaveletsHFBlock
    out = {}
    for i, (C, H, W, B) in enumerate(((8, 6, 10, 2), (16, 4, 4, 3), (4, 12, 2, 1))):
        torch.manual_seed(40 + i)
        m = WaveletsHFBlock(C).eval()
        x = torch.randn(B, C, H, W)
        with torch.no_grad():
            y = m(x)
        out.update({f"x{i}": x.numpy(), f"y{i}": y.numpy(), f"w1_{i}": m.filter1[0].weight.detach().numpy(),
                    f"b1_{i}": m.filter1[0].bias.detach().numpy(), f"w3_{i}": m.filter2[0].weight.detach().numpy(),
                    f"b3_{i}": m.filter2[0].bias.detach().numpy()})
    np.savez_compressed(os.path.join(OUT, "edge.npz"), **out)


if __name__ == "__main__":
    gen_clips()
    gen_repp()
    gen_edge()
    for f in ("clips.json", "repp.json", "edge.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
