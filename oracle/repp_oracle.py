"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float64) of the reference's REPP tubelet post-processing
(tools/REPP.py:24-277 + tools/repp_utils.py:34-109), the step after the aggregation stage in `tscd_demo.py --post`
(:286-290) and the VID evaluation recipe.  Only tests/ may import this; the product is tscd_b200/repp.py + csrc/repp.cu.

Pinned by tests/golden/repp.json: outputs of the reference's own REPP class (run by tools/make_goldens_host.py on seeded
synthetic predictions, `def` and `logreg` distances, with and without re-coordination / unmatched detections).

The learned linking model is a 4-feature logistic regression (tools/matching_model_logreg.pckl: sklearn LogisticRegression,
features center_distances_corrected, height_rel, iou, width_rel); the oracle takes its coefficients as plain numbers
(predict_proba of a binary LogisticRegression = expit(x . coef + intercept))."""
import copy
import math

import numpy as np

INF = 9e15          # REPP.py:16


def get_iou(b1, b2):
    """repp_utils.py:53-109 on (x, y, w, h) boxes."""
    a = [b1[0], b1[1], b1[0] + b1[2], b1[1] + b1[3]]
    b = [b2[0], b2[1], b2[0] + b2[2], b2[1] + b2[3]]
    x_left, y_top = max(a[0], b[0]), max(a[1], b[1])
    x_right, y_bottom = min(a[2], b[2]), min(a[3], b[3])
    if x_right < x_left or y_bottom < y_top:
        return 0.0
    inter = (x_right - x_left) * (y_bottom - y_top)
    return inter / float((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter)


def pair_features(p1, p2, names):
    """repp_utils.py:34-50."""
    f = {}
    if "width_rel" in names:
        f["width_rel"] = min(p1["bbox"][2], p2["bbox"][2]) / max(p1["bbox"][2], p2["bbox"][2])
    if "height_rel" in names:
        f["height_rel"] = min(p1["bbox"][3], p2["bbox"][3]) / max(p1["bbox"][3], p2["bbox"][3])
    if "iou" in names:
        f["iou"] = get_iou(list(p1["bbox"]), list(p2["bbox"]))
    if "center_distances_corrected" in names:
        c1, c2 = p1["bbox_center"], p2["bbox_center"]
        f["center_distances_corrected"] = math.sqrt((c2[0] - c1[0]) ** 2 + (c2[1] - c1[1]) ** 2)
    return f


class REPPOracle:
    def __init__(self, min_tubelet_score, add_unmatched, min_pred_score, distance_func, clf_thr, clf_mode, recoordinate,
                 recoordinate_std, logreg=None, num_classes=30, **_):
        self.min_tubelet_score, self.add_unmatched, self.min_pred_score = min_tubelet_score, add_unmatched, min_pred_score
        self.distance_func, self.clf_thr, self.clf_mode = distance_func, clf_thr, clf_mode
        self.recoordinate, self.recoordinate_std = recoordinate, recoordinate_std
        self.logreg = logreg            # dict(features=[...], coef=[...], intercept=float)
        self.num_classes = num_classes

    # REPP.py:52-57
    def distance_def(self, p1, p2):
        div = get_iou(list(p1["bbox"]), list(p2["bbox"])) * np.dot(p1["scores"], p2["scores"])
        return INF if div == 0 else 1 / div

    # REPP.py:60-79
    def distance_logreg(self, p1, p2):
        f = pair_features(p1, p2, self.logreg["features"])
        x = np.array([[f[c] for c in self.logreg["features"]]])
        d = x @ np.asarray(self.logreg["coef"], dtype=np.float64).reshape(-1, 1) + self.logreg["intercept"]
        score = 1.0 / (1.0 + np.exp(-d[:, 0]))
        if score < self.clf_thr:
            return INF
        if self.clf_mode == "max":
            score = p1["scores"].max() * p2["scores"].max() * score
        elif self.clf_mode == "dot":
            score = np.dot(p1["scores"], p2["scores"]) * score
        elif self.clf_mode == "dot_plus":
            score = np.dot(p1["scores"], p2["scores"]) + score
        elif self.clf_mode != "raw":
            raise ValueError("error post_clf")
        return float((1 - score)[0])

    # REPP.py:116-133 (minimisation branch)
    @staticmethod
    def solve(distances):
        pairs = []
        distances = distances.copy()
        while distances.min() != INF:
            inds = np.where(distances == distances.min())
            a, b = int(inds[0][0]), int(inds[1][0])
            pairs.append((a, b))
            distances[a, :] = INF
            distances[:, b] = INF
        return pairs

    # REPP.py:82-113
    def video_pairs(self, frames, preds):
        match = self.distance_def if self.distance_func == "def" else self.distance_logreg
        pairs, unmatched = [], []
        for i in range(len(frames) - 1):
            p1s, p2s = preds[frames[i]], preds[frames[i + 1]]
            pi = []
            if len(p1s) and len(p2s):
                d = np.zeros((len(p1s), len(p2s)))
                for a, p1 in enumerate(p1s):
                    for b, p2 in enumerate(p2s):
                        d[a, b] = match(p1, p2)
                pi = self.solve(d)
            unmatched.append([a for a in range(len(p1s)) if a not in [p[0] for p in pi]])
            pairs.append(pi)
        return pairs, unmatched

    # REPP.py:138-190
    @staticmethod
    def tubelets(frames, preds, pairs):
        n = len(frames)
        tubs, count, first = [], 0, 0
        while first != n - 1:
            ind = None
            cur = first
            for cur in range(first, n - 1):
                if ind is not None:
                    pr = [p for p in pairs[cur] if p[0] == ind]
                    if not pr:
                        tubs[count].append((cur, preds[frames[cur]][ind]))
                        count += 1
                        ind = None
                        break
                    pr = pr[0]
                    del pairs[cur][pairs[cur].index(pr)]
                    tubs[count].append((cur, preds[frames[cur]][ind]))
                    ind = pr[1]
                else:
                    if not pairs[cur]:
                        first = cur + 1
                        continue
                    pr = pairs[cur][0]
                    del pairs[cur][0]
                    tubs.append([(cur, preds[frames[cur]][pr[0]])])
                    ind = pr[1]
            if ind is not None:
                tubs[count].append((cur + 1, preds[frames[cur + 1]][ind]))
                count += 1
        return tubs

    def __call__(self, video_predictions):
        """REPP.py:245-273.  video_predictions: {frame key: [pred dict(image_id, bbox xywh, bbox_center, scores (obj, cls, id))]}."""
        preds = copy.deepcopy(video_predictions)
        for fr in preds:
            tmp = []
            for p in preds[fr]:
                idx, sc = int(p["scores"][2]), p["scores"][0] * p["scores"][1]
                if sc >= self.min_tubelet_score:
                    p["scores"] = np.zeros([self.num_classes])
                    p["scores"][idx] = sc
                    tmp.append(p)
            preds[fr] = tmp
        frames = sorted(list(preds.keys()), key=int)
        pairs, unmatched = self.video_pairs(frames, preds)
        frames_unsorted = list(preds.keys())        # get_tubelets indexes frames in dict order (REPP.py:141)
        tubs = self.tubelets(frames_unsorted, preds, pairs)
        for t in tubs:                              # rescore_tubelets :193-202
            new = np.mean([p["scores"] for _, p in t], axis=0)
            for _, p in t:
                p["scores"] = new
        if self.recoordinate:                       # recoordinate_tubelets_full :205-218
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
        if self.add_unmatched:                      # :235-243
            lp = list(preds.values())
            for i in range(len(unmatched)):
                for e in unmatched[i]:
                    tubs.append([(i, lp[i][e])])
        out, track = [], 0
        for t in tubs:                              # tubelets_to_predictions :221-233
            for _, p in t:
                for cat, s in enumerate(p["scores"]):
                    if s < self.min_pred_score:
                        continue
                    out.append({"image_id": p["image_id"], "bbox": list(map(float, p["bbox"])), "score": float(s),
                                "category_id": cat, "track_id": track})
            track += 1
        return out
