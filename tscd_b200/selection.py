"""Proposal selection + bank gather (K1-K3) as one host-side call, sync-free.

Mirrors TSCDHead.postprocess_widx / postpro_woclass + find_feature_score (tscd_head.py:1546-1693,
post_process.py:464-521, tscd_head.py:976-1006) but for a whole batch of frames at once and with all
data-dependent sizes left on the device."""
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops


@dataclass
class SelectionConfig:
    mode: str = "B"                 # "A" = postpro_woclass (top-P objectness -> NMS -> first K); "B" = postprocess_widx
    pre_k: int = 750                # mode A: P   (Prenum)
    top_k: int = 30                 # mode A: K   (Afternum)
    nms_thresh: float = 0.75        # pre-NMS IoU threshold (TSCDHead pre_nms)
    conf_thresh: float = 0.001      # mode B
    minimal_limit: int = 0          # mode B
    maximal_limit: int = 0          # mode B
    use_pre_nms: bool = True        # mode B (False in both shipped -L exps)

    def max_keep(self, num_anchors: int) -> int:
        if self.mode == "A":
            return min(self.top_k, self.pre_k, num_anchors)
        cap = self.maximal_limit if self.maximal_limit else num_anchors
        if self.minimal_limit:
            cap = max(cap, min(self.minimal_limit, num_anchors))
        return cap


def select_and_gather(head: ops.HeadViews, feats, feat_dtype, feat_dim, cfg: SelectionConfig,
                      bank_dtype=torch.float16, status: Optional[torch.Tensor] = None, bank_rows: Optional[int] = None):
    """Runs K1 (+K2) + K3.  Returns the dict of ops.gather plus the candidate dict under 'cand'."""
    cand = ops.select(head, cfg.mode, pre_k=cfg.pre_k, conf_thresh=cfg.conf_thresh,
                      minimal_limit=cfg.minimal_limit, maximal_limit=cfg.maximal_limit)
    keep = keep_count = None
    max_keep = cfg.max_keep(head.anchors.num_anchors)
    if cfg.mode == "A" or cfg.use_pre_nms:
        keep, keep_count, status = ops.nms(cand["box"], cand["score"], cand["cls"], cand["count"], cfg.nms_thresh,
                                           max_keep=max_keep, status=status)
    out = ops.gather(head, feats, feat_dtype, feat_dim, cand, keep, keep_count, max_keep=max_keep,
                     bank_dtype=bank_dtype, bank_rows=bank_rows)
    out["cand"] = cand
    out["status"] = status
    return out


def to_lists(sel):
    """Host-side view in the reference's container types: (rows list F x [n,7+C] or None, idx list F x int64[n]).
    This is the only place that synchronises."""
    counts = sel["sel_count"].cpu().tolist()
    rows, idxs = [], []
    for f, n in enumerate(counts):
        if n == 0:
            rows.append(None)
            idxs.append(None)
        else:
            rows.append(sel["sel_rows"][f, :n].clone())
            idxs.append(sel["sel_idx"][f, :n].to(torch.int64))
    return rows, idxs
