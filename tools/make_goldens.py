#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own PyTorch code (fp32, CPU).

Run in the build container only (needs /root/reference):  python tools/make_goldens.py
The fixtures pin ``oracle/`` (tests/test_oracle_golden.py); the GPU box never sees the
reference, only these vectors.  No reference source is copied: the script imports the
reference through tools/ref_shim.py and records inputs/outputs of its functions.

Fixtures (all float32 / int64, seeds fixed):
  select.npz      postprocess_widx (mode B; 4 limit/pre-NMS variants) and postpro_woclass (mode A)
                  on synthetic head outputs, A=1344 anchors (256x256), C=6
  msa.npz         MSA_yolov (gen-1 self-attention aggregation), D=32, N=96
  stage_tscd.npz  TSCDHead.forward inference tail at width 0.125 (D=32): boundary tensors,
                  aggregation-stage weights, every intermediate, final detections; two
                  consecutive clips (resume=False then resume=True) with ragged proposal counts
  thirdparty.npz  torchvision nms / scipy linear_sum_assignment known-answer cases
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

ref_shim.install()

from yolox.models.tscd_head import TSCDHead  # noqa: E402
from yolox.models.post_process import postpro_woclass, postprocess  # noqa: E402
from yolox.models.post_trans import MSA_yolov  # noqa: E402
from yolox.data.datasets.vid import get_timing_signal_1d  # noqa: E402
import torchvision  # noqa: E402
from scipy.optimize import linear_sum_assignment  # noqa: E402

from oracle import tscd_oracle as O  # noqa: E402  (only for the synthetic-input generator)

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

OBJ_MEAN = float(os.environ.get("OBJ_MEAN", "-9.5"))
MORE_ARGS = {'use_ffn': True, 'use_time_emd': False, 'use_loc_emd': True, 'loc_fuse_type': 'identity',
             'use_qkv': True, 'local_mask': False, 'local_mask_branch': '', 'pure_pos_emb': False,
             'loc_conf': False, 'iou_base': False, 'reconf': True, 'ota_mode': True, 'ota_cls': False,
             'traj_linking': False, 'iou_window': 0, 'globalBlocks': 1, 'use_pre_nms': False,
             'cat_ota_fg': False, 'agg_type': 'mca', 'minimal_limit': 50, 'maximal_limit': 500,
             'conf_sim_thresh': 0.99, 'decouple_reg': True}   # exps/TSCD_OVIS/ovis_tscd_large.py:119-129


def npify(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def pack_list(prefix, lst, out):
    """Store a ragged list of tensors (None -> empty marker)."""
    out[prefix + ".len"] = np.int64(len(lst))
    for i, t in enumerate(lst):
        if t is None:
            out[f"{prefix}.{i}.none"] = np.int64(1)
        else:
            out[f"{prefix}.{i}"] = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def make_head(num_classes, width, **over):
    args = dict(MORE_ARGS)
    args.update(over)
    head = TSCDHead(num_classes, width, in_channels=[256, 512, 1024], heads=4, defualt_p=30, defulat_pre=750,
                    pre_nms=0.75, sim_thresh=0.75, ave=True, **args)
    head.initialize_biases(1e-2)
    return head.eval()


# ----------------------------------------------------------------------------- select.npz
def gen_select():
    C = 6
    hw = [(32, 32), (16, 16), (8, 8)]
    head_out, _ = O.synth_head_outputs(4, hw, C, dim=8, seed=2024, clustered=True, obj_mean=[-10.5, -9.0, -7.5, -3.0])
    out = {"head_out": head_out.numpy(), "hw": np.asarray(hw), "strides": np.asarray([8, 16, 32]), "C": np.int64(C)}
    head = make_head(C, 0.125)
    head.hw = [torch.Size(x) for x in hw]
    decoded, _, _ = head.decode_outputs(head_out.clone(), dtype=head_out.type())
    out["decoded"] = decoded.numpy()
    variants = {
        "b_min50_max500": dict(minimal_limit=50, maximal_limit=500, use_pre_nms=False),
        "b_max100": dict(minimal_limit=0, maximal_limit=100, use_pre_nms=False),
        "b_prenms": dict(minimal_limit=0, maximal_limit=0, use_pre_nms=True),
        "b_min700_prenms": dict(minimal_limit=700, maximal_limit=0, use_pre_nms=True),
    }
    for name, kw in variants.items():
        head.kwargs.update(kw)
        rows, idxs, _, _ = head.postprocess_widx(decoded.clone(), num_classes=C, nms_thre=0.75, ota_idxs=None)
        pack_list(name + ".rows", rows, out)
        pack_list(name + ".idx", idxs, out)
        print(name, [None if r is None else r.shape[0] for r in rows])
    rows, idxs = postpro_woclass(decoded.clone(), num_classes=C, nms_thre=0.75, topK=30)
    pack_list("a_750_30.rows", rows, out)
    pack_list("a_750_30.idx", idxs, out)
    print("a_750_30", [r.shape[0] for r in rows])
    # tie check at the top-750 boundary (torch.topk tie order is unspecified)
    for f in range(decoded.shape[0]):
        s = torch.sort(decoded[f, :, 4], descending=True).values
        assert s[749] != s[750], "objectness tie at the top-k boundary; change the seed"
    np.savez_compressed(os.path.join(OUT, "select.npz"), **out)


# ----------------------------------------------------------------------------- msa.npz
def gen_msa():
    torch.manual_seed(7)
    D, N = 32, 96
    m = MSA_yolov(dim=D, out_dim=4 * D, num_heads=4).eval()
    g = torch.Generator().manual_seed(11)
    base = torch.randn(12, D, generator=g)
    # clustered features so the 0.75 cosine mask is non-trivial
    xc = (base[torch.randint(0, 12, (N,), generator=g)] + 0.25 * torch.randn(N, D, generator=g)).unsqueeze(0)
    xr = (base[torch.randint(0, 12, (N,), generator=g)] + 0.25 * torch.randn(N, D, generator=g)).unsqueeze(0)
    cs = torch.rand(N, generator=g)
    fs = torch.rand(N, generator=g)
    with torch.no_grad():
        oc, _ = m(xc, xr, cs, fs, sim_thresh=0.75, ave=True, use_mask=False)
        x_c, x_r, r2c, r2o = m.msa(xc, xr, cs, fs, sim_thresh=0.75, ave=True, use_mask=False)
    out = {"x_cls": xc, "x_reg": xr, "cls_score": cs, "fg_score": fs, "out_cls": oc, "att_x_cls": x_c,
           "att_x_reg": x_r, "r2c": r2c, "r2o": r2o}
    for k, v in m.state_dict().items():
        out["sd.trans." + k] = v
    print("msa: mask density", float((r2c > 0).float().mean()))
    np.savez_compressed(os.path.join(OUT, "msa.npz"), **npify(out))


# ----------------------------------------------------------------------------- stage_tscd.npz
def gen_stage():
    torch.manual_seed(2024)
    C, width = 5, 0.125
    L, G = 3, 3
    F_ = L + G
    head = make_head(C, width, minimal_limit=12, maximal_limit=40)
    # Random-init logits are nearly constant (std ~0.015), so every frame would hit the same limit.
    # Rescale the 1x1 prediction convs (weights only -- reference code untouched) so that the
    # still-detector scores straddle the 0.001 threshold and frames get ragged proposal counts.
    g = torch.Generator().manual_seed(99)
    clips = []
    for clip in range(2):
        clips.append([torch.randn(F_, int(256 * width), 16, 16, generator=g),
                      torch.randn(F_, int(512 * width), 8, 8, generator=g),
                      torch.randn(F_, int(1024 * width), 4, 4, generator=g)])
    with torch.no_grad():
        for k in range(3):
            x = head.stems[k](clips[0][k])
            rf, cf = head.reg_convs[k](x), head.cls_convs[k](x)
            for conv, feat, std, mean in ((head.obj_preds[k], rf, 2.0, OBJ_MEAN), (head.cls_preds[k], cf, 2.0, -1.5),
                                          (head.reg_preds[k], rf, 0.5, 0.3)):
                conv.bias.zero_()
                conv.weight.mul_(std / conv(feat).std())
                conv.bias.fill_(mean - float(conv(feat).mean()))
            head.reg_preds[k].bias[2:] += 1.2            # boxes ~4.5 strides wide -> real overlaps for the final NMS
        # make the refined heads decisive so the final 0.001 filters and NMS(0.5) all do work
        head.cls_pred.weight.mul_(40.0); head.cls_pred.bias.fill_(-4.0)
        head.matcher_obj_pred.weight.mul_(5.0); head.matcher_obj_pred.bias.fill_(-2.0)
    rec = {}

    def hook(name):
        def fn(mod, args, output):
            rec.setdefault(name, []).append(output)
        return fn

    for name in ("agg", "agg_iou", "local_reg_matcher", "fc_reg_matcher", "task_aligned", "cls_pred",
                 "matcher_obj_pred", "matcher_reg_pred"):
        getattr(head, name).register_forward_hook(hook(name))

    orig_ffs, orig_pw, orig_dec = head.find_feature_score, head.postprocess_widx, head.decode_outputs

    def ffs(features, idxs, reg_features, reg_edge_features, imgs=None, predictions=None):
        rec.setdefault("planes", []).append((features.clone(), reg_features.clone(), reg_edge_features.clone()))
        res = orig_ffs(features, idxs, reg_features, reg_edge_features, imgs, predictions)
        rec.setdefault("bank", []).append(res)
        return res

    def pw(prediction, **kw):
        rec.setdefault("decoded", []).append(prediction.clone())
        res = orig_pw(prediction, **kw)
        rec.setdefault("select", []).append((res[0], res[1]))
        return res

    def dec(outputs, dtype, flevel=0):
        rec.setdefault("head_out", []).append(outputs.clone())
        return orig_dec(outputs, dtype, flevel)

    head.find_feature_score, head.postprocess_widx, head.decode_outputs = ffs, pw, dec

    import yolox.models.tscd_matching as tm
    orig_lsa = tm.linear_sum_assignment

    def lsa(cmat):
        r = orig_lsa(cmat)
        rec.setdefault("lap_cost", []).append(np.asarray(cmat, dtype=np.float32).copy())
        rec.setdefault("lap_col", []).append(np.asarray(r[1]).copy())
        return r

    tm.linear_sum_assignment = lsa

    out = {"C": np.int64(C), "L": np.int64(L), "G": np.int64(G), "D": np.int64(int(256 * width)),
           "hw": np.asarray([(16, 16), (8, 8), (4, 4)]), "strides": np.asarray([8, 16, 32])}
    for clip in range(2):
        xin = clips[clip]
        te = get_timing_signal_1d(torch.arange(clip * L, clip * L + L), 256)
        for v in rec.values():
            v.clear()
        with torch.no_grad():
            result, result_ori = head(xin, None, torch.zeros(F_, 3, 128, 128), te, nms_thresh=0.5, lframe=L, gframe=G,
                                      resume=(clip == 1))
        p = f"clip{clip}."
        out[p + "time_embedding"] = te
        out[p + "head_out"] = rec["head_out"][0]
        out[p + "decoded"] = rec["decoded"][0]
        pc, pr, pe = rec["planes"][0]
        out[p + "plane_cls"], out[p + "plane_reg"], out[p + "plane_edge"] = pc.contiguous(), pr.contiguous(), pe.contiguous()
        rows, idxs = rec["select"][0]
        pack_list(p + "rows", rows, out)
        pack_list(p + "idx", idxs, out)
        for j, nm in enumerate(("bank_cls", "bank_reg", "bank_edge", "cls_scores", "fg_scores", "locs", "all_scores")):
            out[p + nm] = rec["bank"][0][j]
        out[p + "agg_cls"] = rec["agg"][0][0]
        out[p + "iou_cls"], out[p + "iou_reg"] = rec["agg_iou"][0]
        out[p + "cafm"] = rec["local_reg_matcher"][0][-1]
        out[p + "fc_reg_matcher"] = rec["fc_reg_matcher"][0]
        out[p + "task_aligned"] = rec["task_aligned"][0]
        out[p + "cls_preds"] = rec["cls_pred"][0]
        out[p + "obj_preds"] = rec["matcher_obj_pred"][0]
        out[p + "reg_deltas"] = rec["matcher_reg_pred"][0]
        pack_list(p + "lap_cost", rec["lap_cost"], out)
        pack_list(p + "lap_col", rec["lap_col"], out)
        pack_list(p + "result", result, out)
        pack_list(p + "result_ori", result_ori, out)
        print("clip", clip, "proposals/frame", [None if i is None else len(i) for i in idxs],
              "dets", [None if r is None else r.shape[0] for r in result],
              "ori", [None if r is None else r.shape[0] for r in result_ori])
    tm.linear_sum_assignment = orig_lsa
    skip = ("stems", "cls_convs", "reg_convs", "edge_enhance", "cls_preds", "reg_preds", "obj_preds")
    for k, v in head.state_dict().items():
        if not k.startswith(skip):
            out["sd." + k] = v
    np.savez_compressed(os.path.join(OUT, "stage_tscd.npz"), **npify(out))


# ----------------------------------------------------------------------------- thirdparty.npz
def gen_thirdparty():
    g = torch.Generator().manual_seed(5)
    out = {}
    # NMS: clustered boxes with class offsets, including near-duplicate scores and negative coordinates
    for t, n in enumerate((1, 7, 64, 300, 750)):
        ctr = torch.rand(max(n // 6, 1), 2, generator=g) * 200 - 20
        which = torch.randint(0, ctr.shape[0], (n,), generator=g)
        c = ctr[which] + torch.randn(n, 2, generator=g) * 4
        wh = torch.rand(n, 2, generator=g) * 40 + 8
        boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
        scores = torch.rand(n, generator=g)
        if n > 10:
            scores[3] = scores[9]            # engineered exact tie -> stable order, lower index first
        cls = torch.randint(0, 4, (n,), generator=g).float()
        for thr in (0.5, 0.75):
            keep = torchvision.ops.batched_nms(boxes, scores, cls, thr)
            out[f"nms{t}.{thr}.keep"] = keep
        out[f"nms{t}.boxes"], out[f"nms{t}.scores"], out[f"nms{t}.cls"] = boxes, scores, cls
    # LAP: square, wide, tall, duplicated rows (ties), constant matrix
    shapes = [(1, 1), (5, 5), (30, 30), (30, 47), (47, 30), (50, 50), (12, 12), (9, 9)]
    for t, (r, c) in enumerate(shapes):
        cost = torch.rand(r, c, generator=g).double()
        if t == 6:
            cost[3] = cost[7]
            cost[:, 2] = cost[:, 5]
        if t == 7:
            cost[:] = 0.25
        a, b = linear_sum_assignment(cost.numpy())
        out[f"lap{t}.cost"], out[f"lap{t}.row"], out[f"lap{t}.col"] = cost, a, b
    np.savez_compressed(os.path.join(OUT, "thirdparty.npz"), **npify(out))


if __name__ == "__main__":
    gen_thirdparty()
    gen_select()
    gen_msa()
    gen_stage()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
