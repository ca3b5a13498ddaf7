"""Runs the UNMODIFIED reference `TSCDHead.forward` (yolox/models/tscd_head.py:303-733) from the hot path's seam.

The aggregation stage starts right after the decoupled-head convolutions.  The reference does not
split its forward there, so this runner replaces the conv-tower *modules* of a stock `TSCDHead` instance (stems, cls/reg
convs, cls/reg/obj preds, edge_enhance_reg -- everything upstream of the seam, out of scope) by replay modules that hand
back prepared tensors, and then calls the stock `forward`: sigmoid + concat + flatten, decode_outputs, postprocess_widx,
find_feature_score, agg / agg_iou (MCA), CAFM with scipy's Hungarian, fc_reg_matcher, TaskAligned, the prediction
Linears, decode_reg_preds5 and postprocess all run as the reference wrote them, on CPU or CUDA.

BASELINE.json configs[1] ("pre-NMS top-750 -> 30 proposals/frame") names the gen-1 selection `postpro_woclass`
(yolox/models/post_process.py:464-521), which TSCDHead never calls (SURVEY finding 1).  `selection="A"` therefore binds the
reference's own, unmodified `postpro_woclass` in place of `postprocess_widx` on the instance -- the one piece of glue here.

No reference source lives in this file; the package comes from /root/reference or from baseline/_ref
(baseline/install_reference.sh).  Used by bench.py --impl reference / cpu_baseline and by the GPU reference tests."""
import os
import sys
import types

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_TOOLS = os.path.join(os.path.dirname(_HERE), "tools")

OVIS_L_ARGS = {'use_ffn': True, 'use_time_emd': False, 'use_loc_emd': True, 'loc_fuse_type': 'identity',
               'use_qkv': True, 'local_mask': False, 'local_mask_branch': '', 'pure_pos_emb': False,
               'loc_conf': False, 'iou_base': False, 'reconf': True, 'ota_mode': True, 'ota_cls': False,
               'traj_linking': False, 'iou_window': 0, 'globalBlocks': 1, 'use_pre_nms': False,
               'cat_ota_fg': False, 'agg_type': 'mca', 'minimal_limit': 50, 'maximal_limit': 500,
               'conf_sim_thresh': 0.99, 'decouple_reg': True}   # values of exps/TSCD_OVIS/ovis_tscd_large.py:32-49,119-129
VID_L_ARGS = {k: v for k, v in OVIS_L_ARGS.items() if k not in ("maximal_limit", "conf_sim_thresh")}   # exps/TSCD_VID/vid_tscd_large.py:39-42,114-123


def available():
    if _TOOLS not in sys.path:
        sys.path.insert(0, _TOOLS)
    import ref_shim
    return ref_shim.reference_root() is not None


def install(cpu_redirect=None):
    if _TOOLS not in sys.path:
        sys.path.insert(0, _TOOLS)
    import ref_shim
    ref_shim.install(cpu_redirect=cpu_redirect)
    return ref_shim


class _Replay(nn.Module):
    """Stands in for a conv-tower module upstream of the seam: returns the tensor prepared for this forward."""

    def __init__(self):
        super().__init__()
        self.value = None

    def forward(self, x):
        return x if self.value is None else self.value


def build_head(num_classes, more_args, seed=2024, state_dict=None):
    """Stock TSCDHead(width 1.0, TSCD-L) in eval mode; random init (seed) or the given aggregation-stage state dict."""
    install()
    from yolox.models.tscd_head import TSCDHead
    torch.manual_seed(seed)
    head = TSCDHead(num_classes, 1.0, in_channels=[256, 512, 1024], heads=4, defualt_p=30, defulat_pre=750,
                    pre_nms=0.75, sim_thresh=0.75, ave=True, **dict(more_args))
    head.initialize_biases(1e-2)
    if state_dict is not None:
        missing = head.load_state_dict(state_dict, strict=False)
        assert not missing.unexpected_keys, missing.unexpected_keys
    return head.eval()


def attach_replay(head, selection="B"):
    """Replace the modules upstream of the seam by replay stubs (in place) and, for selection 'A', bind postpro_woclass."""
    n = len(head.stems)
    for name in ("stems", "cls_convs", "reg_convs", "cls_convs2", "reg_convs2", "cls_preds", "reg_preds", "obj_preds",
                 "edge_enhance_reg"):
        ml = getattr(head, name)
        for k in range(n):
            ml[k] = _Replay()
    if selection == "A":
        from yolox.models.post_process import postpro_woclass

        def _mode_a(self, prediction, num_classes, nms_thre=0.5, ota_idxs=None, conf_thresh=0.001):
            out, idx = postpro_woclass(prediction, num_classes, nms_thre=nms_thre, topK=self.Afternum)
            return out, idx, None, None

        head.postprocess_widx = types.MethodType(_mode_a, head)
    return head


def run_tail(head, reg, obj, cls, f_cls, f_reg, f_edge, time_embedding, lframe, gframe, img_hw=(576, 576), resume=False,
             nms_thresh=0.5):
    """One stock forward from the seam.  reg/obj/cls/f_*: per-level lists of [F, ch, H, W] tensors (raw conv outputs)."""
    for k in range(len(reg)):
        head.reg_preds[k].value, head.obj_preds[k].value, head.cls_preds[k].value = reg[k], obj[k], cls[k]
        head.cls_convs2[k].value, head.reg_convs2[k].value, head.edge_enhance_reg[k].value = f_cls[k], f_reg[k], f_edge[k]
    F = reg[0].shape[0]
    imgs = types.SimpleNamespace(shape=(F, 3, img_hw[0], img_hw[1]))          # forward only reads imgs.shape
    xin = [reg[k] for k in range(len(reg))]                                   # xin[0].type() / dtype are read; the stubs ignore it
    with torch.no_grad():
        return head(xin, None, imgs, time_embedding, nms_thresh=nms_thresh, lframe=lframe, gframe=gframe, resume=resume)
