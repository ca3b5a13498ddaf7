"""Gen-1 (YOLOV) aggregation stage named by the north_star items (1)-(4): postpro_woclass (top-P objectness -> class-aware
NMS -> first K; yolox/models/post_process.py:464-521) -> feature gather -> MSA_yolov self-attention over ALL F x K proposals
of a clip (post_trans.py:717-826, 1227-1269) -> linear_pred (yolovp_msa.py:326-376).  The reference's own gen-1 head cannot
finish its forward as shipped (SURVEY finding 3: `postprocess` is called without its required `reg_output`), so -- like the
oracle's stage_gen1 -- the stage ends at the refined class logits.  Same kernels as the TSCD stage (self_attn layout)."""
from typing import Dict

import torch

from . import _lib as L
from . import aggregate, ops
from . import selection as _selection
from .selection import SelectionConfig
from .stage import MAX_KEYS_PER_CLIP, _r128


class Gen1Stage:
    def __init__(self, num_classes: int, selection: SelectionConfig, state_dict: Dict[str, torch.Tensor], device="cuda",
                 dtype=torch.float16, sim_thresh: float = 0.75):
        L.lib()
        if selection.mode != "A":
            raise RuntimeError("the gen-1 stage selects with postpro_woclass (mode 'A')")
        selection.validate(chain_capacity=4096)
        self.num_classes, self.sel, self.device, self.dtype, self.sim_thresh = num_classes, selection, device, dtype, sim_thresh
        self.w = aggregate.MSAWeights(state_dict, "trans.", dtype, device)
        self.pred_w = state_dict["linear_pred.weight"].detach().to(device=device, dtype=dtype).contiguous()
        self.pred_b = state_dict["linear_pred.bias"].detach().to(device=device, dtype=torch.float32).contiguous()

    def forward(self, head: ops.HeadViews, feats, feat_dtype, B: int, F: int):
        """head / feats: seam tensors of B clips x F frames (feats = (cls, reg, reg) views: gen-1 has no edge branch).
        Returns dict(sel, layout, msa16, msa32, logits [row_cap, C+1] fp32, status); rows of clip b are
        [row_off[b*F], row_off[(b+1)*F])."""
        dev, dt = self.device, self.dtype
        kmax = self.sel.max_keep(head.anchors.num_anchors)
        if _r128(F * kmax) > MAX_KEYS_PER_CLIP:
            raise RuntimeError(f"{F} frames x {kmax} proposals exceed {MAX_KEYS_PER_CLIP} keys per clip")
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        row_cap = _r128(B * F * kmax) + 128
        sel = _selection.select_and_gather(head, feats, feat_dtype, 256, self.sel, bank_dtype=dt, status=status, bank_rows=row_cap)
        lay = aggregate.make_layout(sel["sel_count"], B, F, F, row_cap, row_cap, _r128(F * kmax), dt, row_off=sel["row_off"])
        lay.self_attn = True
        lay.lrow_off = lay.row_off
        n_rows = lay.row_off[-1:]
        msa16, msa32 = aggregate.msa_forward(lay, self.w, sel["bank_cls"], sel["bank_reg"], sel["bank_score"], n_rows,
                                             sim_thresh=self.sim_thresh)
        _, logits = ops.linear(msa16, self.pred_w, self.pred_b, m_dev=n_rows, want16=False, want32=True, tag="linear_pred")
        return dict(sel=sel, layout=lay, msa16=msa16, msa32=msa32, logits=logits, status=status)
