"""Evaluator / Predictor glue: the consumers of the stage's detections.

The reference turns every frame's `Tensor[n,7]` into Python dicts with a per-frame `.cpu()` and a per-box loop
(OVISEvaluator.convert_to_coco_format, yolox/evaluators/ovis_evaluator_v2.py:233-289; Predictor.to_repp_heavy,
tools/val_to_imdb.py:193-218).  Here all detections of all local frames of a batch are packed on the device into one table
(tscd_pack_detections: resize scale applied, xyxy -> top-left + width/height, score = obj * cls), copied to the host ONCE, and the
dicts are built from plain Python lists -- same keys, same values, same order as the reference's loops."""
from typing import List, Optional, Sequence

import torch

from . import _lib as L
from . import ops


def pack_detections(rows: torch.Tensor, count: torch.Tensor, scales: Optional[torch.Tensor] = None):
    """rows [F, cap, 7] (stage output `det_rows` / `ori_rows`), count [F] int32, scales [F] fp32 (boxes are divided by it).
    Returns (packed [F*cap, 12] device tensor, offsets [F+1] int32 device tensor); nothing is synchronised."""
    Fn, cap, w = rows.shape
    assert w == 7 and rows.dtype == torch.float32 and rows.is_contiguous() and count.dtype == torch.int32
    packed = torch.empty(Fn * cap, 12, dtype=torch.float32, device=rows.device)
    offsets = torch.empty(Fn + 1, dtype=torch.int32, device=rows.device)
    if scales is not None:
        scales = scales.to(device=rows.device, dtype=torch.float32).contiguous()
    ops.call("tscd_pack_detections", L.PackDetectionsArgs, num_frames=Fn, cap=cap, rows=rows, count=count, scale=scales,
             offsets=offsets, packed=packed)
    return packed, offsets


def to_host(packed: torch.Tensor, offsets: torch.Tensor):
    """ONE device->host copy of the valid rows (+ the offsets).  Returns (numpy [N,12], offsets list)."""
    off = offsets.cpu().tolist()
    return packed[:off[-1]].cpu().numpy(), off


def coco_predictions(table, offsets: Sequence[int], det_cand: Sequence[int], first_image_id: int = 0) -> List[dict]:
    """The `data_list` of convert_to_coco_format (ovis_evaluator_v2.py:268-287): one dict per detection, image ids counted per frame
    (frames without detections still consume an id).  det_cand[f] == 0 marks frames whose reference output is None."""
    out = []
    for f in range(len(offsets) - 1):
        if det_cand[f] == 0:
            continue
        for i in range(offsets[f], offsets[f + 1]):
            r = table[i]
            out.append({"image_id": int(first_image_id + f), "category_id": int(r[6]), "bbox": [float(r[1]), float(r[2]), float(r[3]), float(r[4])],
                        "score": float(r[5]), "segmentation": []})
    return out


def repp_predictions(table, offsets: Sequence[int], det_cand: Sequence[int], img_sizes: Sequence[Sequence[int]], image_ids: Sequence) -> List[List[dict]]:
    """Predictor.to_repp_heavy (tools/val_to_imdb.py:193-218) per frame: clipped top-left / width / height boxes, the normalised
    box centre REPP links on, and the (obj, cls_score, class) triple as `scores`.  The table must have been packed with the
    frames' resize ratios as scales."""
    res = []
    for f in range(len(offsets) - 1):
        preds = []
        if det_cand[f] != 0:
            ih, iw = img_sizes[f]
            width_diff, height_diff = max(0, (ih - iw) // 2), max(0, (iw - ih) // 2)
            for i in range(offsets[f], offsets[f + 1]):
                r = table[i]
                x_min, y_min = max(0, r[1]), max(0, r[2])
                x_max, y_max = min(iw, r[8]), min(ih, r[9])
                width, height = x_max - x_min, y_max - y_min
                if width <= 0 or height <= 0:
                    continue
                preds.append({"image_id": image_ids[f], "bbox": [x_min, y_min, width, height],
                              "bbox_center": [(x_min + width_diff + width / 2) / max(iw, ih), (y_min + height_diff + height / 2) / max(iw, ih)],
                              "scores": [r[7], r[10], r[6]]})
        res.append(preds)
    return res
