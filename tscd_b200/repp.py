"""REPP tubelet post-processing: the step after the aggregation stage in `tscd_demo.py --post`
(tools/tscd_demo.py:286-290) and the VID evaluation recipe.

`REPPB200` keeps the constructor arguments and the `__call__(video_predictions) -> predictions_coco` contract of the
reference's `REPP` class (tools/REPP.py:24-277).  The part that dominates the reference -- an n1 x n2 Python double loop per
pair of consecutive frames with one sklearn `predict_proba` call per detection pair, followed by a repeated global arg-min --
runs on the device for ALL frame pairs of the video at once (csrc/repp.cu tscd_repp_link, one CTA per frame pair).  What remains
on the host is the reference's own sequential bookkeeping over the linked detections (tubelet chains, score averaging, the
Gaussian re-coordination of box tracks, the output dicts): O(detections), pointer chasing, no arithmetic worth a kernel.

The logistic linking model is read from the reference's pickle (tools/matching_model_logreg.pckl, an sklearn
LogisticRegression over [center_distances_corrected, height_rel, iou, width_rel]) when `weight_path` is given, or passed as
plain numbers (`logreg=dict(features=..., coef=..., intercept=...)`) -- the kernel needs the four coefficients only.
Appearance matching (`descriptor_dist`, needs per-detection embeddings the TSCD tools never produce: `appearance_matching`
is false in tools/yolo_repp_cfg.json) is rejected loudly."""
import pickle
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from . import ops

_FEATURES = ["center_distances_corrected", "height_rel", "iou", "width_rel"]
_CLF_MODES = {"dot": 0, "max": 1, "dot_plus": 2, "raw": 3}


class REPPB200:
    def __init__(self, min_tubelet_score, add_unmatched, min_pred_score, distance_func, clf_thr, clf_mode,
                 appearance_matching=False, recoordinate=True, recoordinate_std=0.4, weight_path=None, logreg: Optional[dict] = None,
                 num_classes: int = 30, device="cuda"):
        L.lib()                                   # no CPU fallback for the linking step
        if appearance_matching:
            raise RuntimeError("REPPB200: appearance_matching needs detection embeddings; not produced by the TSCD tools")
        if distance_func not in ("def", "logreg"):
            raise ValueError(f"distance_func not recognized: {distance_func}")
        self.min_tubelet_score, self.add_unmatched, self.min_pred_score = min_tubelet_score, add_unmatched, min_pred_score
        self.distance_func, self.clf_thr, self.clf_mode = distance_func, float(clf_thr), clf_mode
        self.recoordinate, self.recoordinate_std = recoordinate, recoordinate_std
        self.num_classes, self.device = num_classes, device
        self.coef, self.intercept = [0.0] * 4, 0.0
        if distance_func == "logreg":
            if clf_mode not in _CLF_MODES:
                raise ValueError("error post_clf")
            if logreg is None:
                if weight_path is None:
                    raise ValueError("REPPB200: distance_func='logreg' needs weight_path or logreg")
                model, feats = pickle.load(open(weight_path, "rb"))
                logreg = dict(features=list(feats), coef=np.asarray(model.coef_).reshape(-1).tolist(),
                              intercept=float(np.asarray(model.intercept_).reshape(-1)[0]))
            if list(logreg["features"]) != _FEATURES:
                raise RuntimeError(f"REPPB200: the linking model must use the features {_FEATURES}, got {logreg['features']}")
            self.coef, self.intercept = [float(c) for c in logreg["coef"]], float(logreg["intercept"])

    # ------------------------------------------------------------------------------------------------------------
    def link(self, frame_off: List[int], bbox: np.ndarray, center: np.ndarray, score: np.ndarray, cls: np.ndarray):
        """tscd_repp_link for one video.  Returns per frame pair the list of (a, b) pairs in extraction order."""
        nf = len(frame_off) - 1
        if nf < 2:
            return [[] for _ in range(max(0, nf - 1))]
        counts = [frame_off[i + 1] - frame_off[i] for i in range(nf)]
        max_det = max(1, max(counts))
        ws_pitch = max(1, max(counts[i] * counts[i + 1] for i in range(nf - 1)))
        dev = self.device
        t_off = torch.tensor(frame_off, dtype=torch.int32, device=dev)
        t_box = torch.from_numpy(np.ascontiguousarray(bbox, dtype=np.float32)).to(dev)
        t_ctr = torch.from_numpy(np.ascontiguousarray(center, dtype=np.float32)).to(dev)
        t_sc = torch.from_numpy(np.ascontiguousarray(score, dtype=np.float64)).to(dev)
        t_cls = torch.from_numpy(np.ascontiguousarray(cls, dtype=np.int32)).to(dev)
        pairs = torch.empty(nf - 1, max_det, 2, dtype=torch.int32, device=dev)
        pair_count = torch.empty(nf - 1, dtype=torch.int32, device=dev)
        ws_dist = torch.empty(nf - 1, ws_pitch, dtype=torch.float64, device=dev)
        ws_idx = torch.empty(nf - 1, ws_pitch, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        a = L.ReppLinkArgs()
        a.num_frames, a.max_det = nf, max_det
        a.distance_func = 0 if self.distance_func == "def" else 1
        a.clf_mode = _CLF_MODES.get(self.clf_mode, 0)
        a.clf_thr, a.intercept = self.clf_thr, self.intercept
        for i in range(4):
            a.coef[i] = self.coef[i]
        a.frame_off, a.bbox, a.center, a.score, a.cls = ops._p(t_off), ops._p(t_box), ops._p(t_ctr), ops._p(t_sc), ops._p(t_cls)
        a.pairs, a.pair_count, a.ws_dist, a.ws_idx, a.ws_pitch, a.status = ops._p(pairs), ops._p(pair_count), ops._p(ws_dist), ops._p(ws_idx), ws_pitch, ops._p(status)
        import ctypes as C
        with L.timed("tscd_repp_link"):
            L.check(L.lib().tscd_repp_link(C.byref(a), ops._stream()), "tscd_repp_link")
        pc = pair_count.cpu().tolist()
        if int(status.item()) != 0:
            raise RuntimeError("tscd_repp_link: more than 4096 detections in a frame")
        ph = pairs.cpu().numpy()
        return [[(int(ph[f, k, 0]), int(ph[f, k, 1])) for k in range(pc[f])] for f in range(nf - 1)]

    # ------------------------------------------------------------------------------------------------------------
    def __call__(self, video_predictions: Dict):
        """REPP.__call__ (tools/REPP.py:245-273).  video_predictions: {frame key: [dict(image_id, bbox [x,y,w,h], bbox_center,
        scores (obj, cls_score, class id))]} as Predictor.to_repp_heavy builds it (tools/val_to_imdb.py:193-218)."""
        preds = {}
        for fr, plist in video_predictions.items():               # :247-256 low-score filter, one-hot score vectors
            tmp = []
            for p in plist:
                idx, sc = int(p["scores"][2]), p["scores"][0] * p["scores"][1]
                if sc >= self.min_tubelet_score:
                    q = dict(p)
                    q["scores"] = np.zeros([self.num_classes])
                    q["scores"][idx] = sc
                    q["_cls"], q["_score"] = idx, float(sc)
                    tmp.append(q)
            preds[fr] = tmp
        frames = sorted(list(preds.keys()), key=int)               # get_video_pairs :84-85
        off, boxes, ctrs, scs, cls = [0], [], [], [], []
        for fr in frames:
            for p in preds[fr]:
                boxes.append([float(np.float32(v)) for v in p["bbox"]])
                ctrs.append([float(np.float32(v)) for v in p["bbox_center"]])
                scs.append(p["_score"])
                cls.append(p["_cls"])
            off.append(len(boxes))
        pairs = self.link(off, np.asarray(boxes, dtype=np.float32).reshape(-1, 4), np.asarray(ctrs, dtype=np.float32).reshape(-1, 2),
                          np.asarray(scs, dtype=np.float64), np.asarray(cls, dtype=np.int32))
        unmatched = []
        for i in range(len(frames) - 1):                          # :109
            linked = {a for a, _ in pairs[i]}
            unmatched.append([a for a in range(len(preds[frames[i]])) if a not in linked])
        tubs = self._tubelets(list(preds.keys()), preds, pairs)
        for t in tubs:                                             # rescore_tubelets :193-202
            new = np.mean([p["scores"] for _, p in t], axis=0)
            for _, p in t:
                p["scores"] = new
        if self.recoordinate:                                      # recoordinate_tubelets_full :205-218
            from scipy import ndimage
            from scipy.signal import windows
            for t in tubs:
                c = np.array([p["bbox"] for _, p in t])
                w = windows.gaussian(len(c) * 2 - 1, std=self.recoordinate_std * 100 / 40)
                w /= sum(w)
                for k in range(4):
                    c[:, k] = ndimage.convolve(c[:, k], w, mode="reflect")
                for j, (_, p) in enumerate(t):
                    p["bbox"] = c[j, :].tolist()
        if self.add_unmatched:                                     # :235-243
            lp = list(preds.values())
            for i in range(len(unmatched)):
                for e in unmatched[i]:
                    tubs.append([(i, lp[i][e])])
        out, track = [], 0
        for t in tubs:                                             # tubelets_to_predictions :221-233
            for _, p in t:
                for cat, s in enumerate(p["scores"]):
                    if s < self.min_pred_score:
                        continue
                    out.append({"image_id": p["image_id"], "bbox": list(map(float, p["bbox"])), "score": float(s),
                                "category_id": cat, "track_id": track})
            track += 1
        return out

    @staticmethod
    def _tubelets(frames, preds, pairs):
        """get_tubelets (tools/REPP.py:138-190): follow the links frame to frame; a new tubelet starts at the first unused
        pair of the earliest frame that still has one."""
        pairs = [list(p) for p in pairs]
        n = len(frames)
        tubs, count, first = [], 0, 0
        while first != n - 1 and n > 1:
            ind = None
            cur = first
            for cur in range(first, n - 1):
                if ind is not None:
                    nxt = next((p for p in pairs[cur] if p[0] == ind), None)
                    if nxt is None:                                # tubelet ended
                        tubs[count].append((cur, preds[frames[cur]][ind]))
                        count += 1
                        ind = None
                        break
                    pairs[cur].remove(nxt)
                    tubs[count].append((cur, preds[frames[cur]][ind]))
                    ind = nxt[1]
                else:
                    if not pairs[cur]:
                        first = cur + 1
                        continue
                    nxt = pairs[cur].pop(0)
                    tubs.append([(cur, preds[frames[cur]][nxt[0]])])
                    ind = nxt[1]
            if ind is not None:                                    # finished in the last frame
                tubs[count].append((cur + 1, preds[frames[cur + 1]][ind]))
                count += 1
        return tubs
