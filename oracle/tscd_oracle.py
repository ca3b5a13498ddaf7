"""CPU restatement (torch fp32 + plain C) of the TSCD aggregation stage.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites the
reference file:line (relative to the reference checkout) whose arithmetic it
follows, expression order included, so that results agree with the reference's
fp32 CPU path to rounding noise (selection / NMS / assignment indices: exactly).

Weights are passed as a flat ``dict`` with the reference's ``state_dict`` key
names, e.g. ``sd['agg.mca.kv_cls.weight']``.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "build_c", "timing_signal_1d", "anchor_grid", "decode_outputs", "cxcywh_to_xyxy",
    "nms", "batched_nms", "topk_lower_index_first", "select_mode_b", "select_mode_a",
    "find_feature_score", "attention_mca_g2l", "mca_tscd_g2l_reg", "attention_msa", "msa_yolov",
    "lap", "se_module", "l2norm_attention", "CAFMState", "aware_position_reg_matcher",
    "task_aligned", "decode_reg_preds5", "postprocess", "stage_tscd", "stage_gen1",
    "synth_head_outputs", "init_stage_weights",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force=False):
    """Compile oracle_kernels.c -> oracle/_build/liboracle.so with gcc (no FMA contraction)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(_HERE, "oracle_kernels.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_nms.restype = ctypes.c_int64
        lib.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p]
        lib.oracle_lap.restype = ctypes.c_int
        lib.oracle_lap.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = lib
    return _LIB


# --------------------------------------------------------------------------- a13
def timing_signal_1d(index_sequence, channels, min_timescale=1.0, max_timescale=1.0e4):
    """yolox/data/datasets/vid.py:1015-1023 (dup tools/tscd_demo.py:152-166)."""
    num_timescales = channels // 2
    log_inc = torch.tensor(math.log(max_timescale / min_timescale) / (num_timescales - 1))
    inv = min_timescale * torch.exp(torch.arange(0, num_timescales) * -log_inc)
    scaled = torch.unsqueeze(index_sequence, 1) * torch.unsqueeze(inv, 0)
    return torch.cat([torch.sin(scaled), torch.cos(scaled)], dim=1)


# --------------------------------------------------------------------------- a3
def anchor_grid(hw, strides):
    """Level-major, row-major (x, y) grid and per-anchor stride; tscd_head.py:756-766."""
    grids, st = [], []
    for (h, w), s in zip(hw, strides):
        yv, xv = torch.meshgrid([torch.arange(h), torch.arange(w)], indexing="ij")
        g = torch.stack((xv, yv), 2).view(1, -1, 2)
        grids.append(g)
        st.append(torch.full((1, g.shape[1], 1), s))
    return torch.cat(grids, 1).float(), torch.cat(st, 1).float()


def decode_outputs(outputs, hw, strides):
    """tscd_head.py:755-770.  outputs [F,A,5+C] (reg | sigma(obj) | sigma(cls)) -> cxcywh, new tensor."""
    grids, st = anchor_grid(hw, strides)
    out = outputs.clone()
    out[..., :2] = (outputs[..., :2] + grids) * st
    out[..., 2:4] = torch.exp(outputs[..., 2:4]) * st
    return out


def cxcywh_to_xyxy(pred):
    """tscd_head.py:1561-1566 / post_process.py:476-481 (divide by 2, then add/sub)."""
    out = pred.clone()
    out[..., 0] = pred[..., 0] - pred[..., 2] / 2
    out[..., 1] = pred[..., 1] - pred[..., 3] / 2
    out[..., 2] = pred[..., 0] + pred[..., 2] / 2
    out[..., 3] = pred[..., 1] + pred[..., 3] / 2
    return out


# --------------------------------------------------------------------------- a14
def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms CPU semantics (plain-C restatement in oracle_kernels.c)."""
    b = np.ascontiguousarray(boxes.detach().cpu().numpy(), dtype=np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(scores.detach().cpu().numpy(), dtype=np.float32).reshape(-1)
    n = b.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    nk = _lib().oracle_nms(b.ctypes.data, s.ctypes.data, n, float(iou_threshold), keep.ctypes.data)
    return torch.from_numpy(keep[:nk].copy())


def batched_nms(boxes, scores, idxs, iou_threshold):
    """torchvision/ops/boxes.py `_batched_nms_coordinate_trick` (the path the reference
    takes on CUDA for every size this stage produces): offsets = idxs * (max + 1)."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
    return nms(boxes + offsets[:, None], scores, iou_threshold)


def topk_lower_index_first(values, k):
    """torch.topk(sorted=True) with the documented tie-break: equal values -> lower index first
    (torch leaves tie order unspecified; tscd_head.py:1598,1605, post_process.py:507)."""
    order = torch.sort(values, descending=True, stable=True).indices
    return order[:k]


# --------------------------------------------------------------------------- a4-B
def _rows(image_pred, num_classes):
    """tscd_head.py:1579-1582 / post_process.py:500-504: [box4,obj,class_conf,class_pred,cls*C]."""
    class_conf, class_pred = torch.max(image_pred[:, 5:5 + num_classes], 1, keepdim=True)
    return torch.cat((image_pred[:, :5], class_conf, class_pred.float(), image_pred[:, 5:5 + num_classes]), 1)


def select_mode_b(decoded, num_classes, nms_thre=0.75, conf_thresh=0.001, minimal_limit=0, maximal_limit=0,
                  use_pre_nms=True):
    """TSCDHead.postprocess_widx, inference branch (ota_idxs=None); tscd_head.py:1546-1693.

    decoded: [F,A,5+C] cxcywh (output of decode_outputs).  Returns (rows list F x [n,7+C],
    idx list F x int64[n]); a frame with no survivor yields (None, None)."""
    pred = cxcywh_to_xyxy(decoded)
    rows_out, idx_out = [], []
    for image_pred in pred:
        det = _rows(image_pred, num_classes)
        score = det[:, 4] * det[:, 5]
        mask = score >= conf_thresh
        if minimal_limit != 0 and int(mask.sum()) < minimal_limit:          # :1594-1599
            mask[topk_lower_index_first(score, minimal_limit)] = True
        if maximal_limit != 0 and int(mask.sum()) > maximal_limit:          # :1600-1607
            top = topk_lower_index_first(score, maximal_limit)
            mask = torch.zeros_like(mask)
            mask[top] = True
        conf_idx = torch.where(mask)[0]                                     # :1622 ascending anchor order
        det = det[mask]
        if det.shape[0] == 0:
            rows_out.append(None)
            idx_out.append(None)
            continue
        if use_pre_nms:                                                     # :1629-1635
            keep = batched_nms(det[:, :4], det[:, 4] * det[:, 5], det[:, 6], nms_thre)
        else:
            keep = torch.arange(det.shape[0])
        rows_out.append(det[keep])
        idx_out.append(conf_idx[keep])
    return rows_out, idx_out


# --------------------------------------------------------------------------- a4-A
def select_mode_a(decoded, num_classes, nms_thre=0.75, pre_k=750, top_k=30):
    """postpro_woclass, post_process.py:464-521 (k=750 hard-coded there; the parametrised twin is
    yolovp_msa.py:920-978 with self.Prenum).  top-P by objectness, class-aware NMS, first K."""
    pred = cxcywh_to_xyxy(decoded)
    rows_out, idx_out = [], []
    for image_pred in pred:
        det = _rows(image_pred, num_classes)
        sort_idx = topk_lower_index_first(image_pred[:, 4], min(pre_k, image_pred.shape[0]))
        tmp = det[sort_idx]
        keep = batched_nms(tmp[:, :4], tmp[:, 4] * tmp[:, 5], tmp[:, 6], nms_thre)
        topk_idx = sort_idx[keep[:top_k]]
        rows_out.append(det[topk_idx])
        idx_out.append(topk_idx)
    return rows_out, idx_out


# --------------------------------------------------------------------------- a5
def find_feature_score(cls_feat, reg_feat, edge_feat, idxs, rows, num_classes):
    """tscd_head.py:976-1006.  feature planes [F,A,D]; returns the clip bank."""
    fc, fr, fe, cs, fg, locs, alls = [], [], [], [], [], [], []
    for i in range(cls_feat.shape[0]):
        if idxs[i] is None or len(idxs[i]) == 0:
            continue
        fc.append(cls_feat[i, idxs[i]])
        fr.append(reg_feat[i, idxs[i]])
        fe.append(edge_feat[i, idxs[i]])
        cs.append(rows[i][:, 5])
        fg.append(rows[i][:, 4])
        locs.append(rows[i][:, :4])
        alls.append(rows[i][:, -num_classes:])
    if not fc:
        return None
    return (torch.cat(fc), torch.cat(fr), torch.cat(fe), torch.cat(cs), torch.cat(fg), torch.cat(locs),
            torch.cat(alls))


# --------------------------------------------------------------------------- a7
def _round2(attn, raw_cls, raw_reg, num_heads, sim_thresh, conf_sim_thresh):
    """post_trans.py:692-709 (same in Attention_msa :803-824): the `ave` round."""
    raw_cls = torch.sum(raw_cls, dim=1)[0] / num_heads
    raw_reg = torch.sum(raw_reg, dim=1)[0] / num_heads
    sim_mask = (raw_cls > sim_thresh).to(attn.dtype)
    obj_mask = (raw_reg > conf_sim_thresh).to(attn.dtype)
    sim_attn = torch.sum(attn, dim=1)[0] / num_heads
    r2 = torch.softmax(sim_attn, dim=-1)
    r2 = sim_mask * r2 / torch.sum(sim_mask * r2, dim=-1, keepdim=True)
    obj = obj_mask * r2 / torch.sum(obj_mask * r2, dim=-1, keepdim=True)
    return r2, obj


def attention_mca_g2l(sd, prefix, x_cls, x_reg, cls_score, n_local, num_heads=4, scale=25.0, sim_thresh=0.75,
                      conf_sim_thresh=0.99):
    """Attention_mca_g2l.forward, post_trans.py:601-714 (reconf=True, ave=True, use_mask=False).

    x_cls, x_reg: [1,N2,D], first n_local rows are the key-frame (query) rows."""
    B, N2, C = x_cls.shape
    N1, H = n_local, num_heads
    q_cls = F.linear(x_cls[:, :N1], sd[prefix + "q_cls_local.weight"]).reshape(B, N1, 1, H, C // H).permute(2, 0, 3, 1, 4)[0]
    kv_cls = F.linear(x_cls, sd[prefix + "kv_cls.weight"]).reshape(B, N2, 2, H, C // H).permute(2, 0, 3, 1, 4)
    q_reg = F.linear(x_reg[:, :N1], sd[prefix + "q_reg_local.weight"]).reshape(B, N1, 1, H, C // H).permute(2, 0, 3, 1, 4)[0]
    kv_reg = F.linear(x_reg, sd[prefix + "kv_reg.weight"]).reshape(B, N2, 2, H, C // H).permute(2, 0, 3, 1, 4)
    k_cls, v_cls, k_reg, v_reg = kv_cls[0], kv_cls[1], kv_reg[0], kv_reg[1]
    q_cls = q_cls / torch.norm(q_cls, dim=-1, keepdim=True)                  # :623-628, no epsilon
    k_cls = k_cls / torch.norm(k_cls, dim=-1, keepdim=True)
    q_reg = q_reg / torch.norm(q_reg, dim=-1, keepdim=True)
    k_reg = k_reg / torch.norm(k_reg, dim=-1, keepdim=True)
    v_cls_n = v_cls / torch.norm(v_cls, dim=-1, keepdim=True)
    v_reg_n = v_reg / torch.norm(v_reg, dim=-1, keepdim=True)
    score = torch.reshape(cls_score, [1, 1, 1, -1]).repeat(1, H, N1, 1)
    raw_cls = v_cls_n[:, :, :N1, :] @ v_cls_n.transpose(-2, -1)              # :642-643
    raw_reg = v_reg_n[:, :, :N1, :] @ v_reg_n.transpose(-2, -1)
    attn_cls = ((q_cls @ k_cls.transpose(-2, -1)) * scale * score).softmax(dim=-1)   # :658,672
    attn_reg = ((q_reg @ k_reg.transpose(-2, -1)) * scale).softmax(dim=-1)           # :660,675
    attn = (attn_reg + attn_cls) / 2                                         # :678
    x = (attn @ v_cls).transpose(1, 2).reshape(B, N1, C)
    x_c = torch.cat([x, v_cls[:, :, :N1, :].permute(0, 2, 1, 3).reshape(B, N1, C)], dim=-1)
    xr = (attn @ v_reg).transpose(1, 2).reshape(B, N1, C)
    x_r = torch.cat([xr, v_reg[:, :, :N1, :].permute(0, 2, 1, 3).reshape(B, N1, C)], dim=-1)
    x_c = F.linear(x_c, sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])           # :687
    x_r = F.linear(x_r, sd[prefix + "linear_reg.weight"], sd[prefix + "linear_reg.bias"])   # :689
    r2c, r2o = _round2(attn, raw_cls, raw_reg, H, sim_thresh, conf_sim_thresh)
    V_cls = v_cls.permute(0, 2, 1, 3).reshape(B, N2, C)[0]
    V_reg = v_reg.permute(0, 2, 1, 3).reshape(B, N2, C)[0]
    out_cls = torch.cat([r2c @ V_cls, x_c[0]], dim=-1)                       # :581-599
    out_reg = torch.cat([r2o @ V_reg, x_r[0]], dim=-1)
    return out_cls, out_reg


def mca_tscd_g2l_reg(sd, prefix, x_cls, x_reg, cls_score, preds_per_frame, lframe, num_heads=4, sim_thresh=0.75,
                     conf_sim_thresh=0.99):
    """MCA_tscd_g2l_reg.forward, post_trans.py:1127-1162, as written (re-projects the global bank
    once per local frame).  x_cls, x_reg: [1,N,D]; returns ([Nloc,4D],[Nloc,4D])."""
    n_loc = sum(preds_per_frame[:lframe])
    xg_cls, xg_reg, sg = x_cls[:, n_loc:], x_reg[:, n_loc:], cls_score[n_loc:]
    start, outs_c, outs_r = 0, [], []
    for i in range(lframe):
        n = preds_per_frame[i]
        xc = torch.cat((x_cls[:, start:start + n], xg_cls), dim=1)
        xr = torch.cat((x_reg[:, start:start + n], xg_reg), dim=1)
        sc = torch.cat((cls_score[start:start + n], sg), dim=0)
        c, r = attention_mca_g2l(sd, prefix + "mca.", xc, xr, sc, n, num_heads, 25.0, sim_thresh, conf_sim_thresh)
        outs_c.append(c)
        outs_r.append(r)
        start += n
    tc = F.linear(torch.cat(outs_c, 0), sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])
    to = F.linear(torch.cat(outs_r, 0), sd[prefix + "linear_obj.weight"], sd[prefix + "linear_obj.bias"])
    return tc, to


# --------------------------------------------------------------------------- a7'
def attention_msa(sd, prefix, x_cls, x_reg, cls_score, num_heads=4, scale=25.0, sim_thresh=0.75,
                  conf_sim_thresh=0.99):
    """Attention_msa.forward, post_trans.py:734-826 (ave=True, use_mask=False)."""
    B, N, C = x_cls.shape
    H = num_heads
    qkv_cls = F.linear(x_cls, sd[prefix + "qkv_cls.weight"]).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    qkv_reg = F.linear(x_reg, sd[prefix + "qkv_reg.weight"]).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q_cls, k_cls, v_cls = qkv_cls[0], qkv_cls[1], qkv_cls[2]
    q_reg, k_reg, v_reg = qkv_reg[0], qkv_reg[1], qkv_reg[2]
    q_cls = q_cls / torch.norm(q_cls, dim=-1, keepdim=True)
    k_cls = k_cls / torch.norm(k_cls, dim=-1, keepdim=True)
    q_reg = q_reg / torch.norm(q_reg, dim=-1, keepdim=True)
    k_reg = k_reg / torch.norm(k_reg, dim=-1, keepdim=True)
    v_cls_n = v_cls / torch.norm(v_cls, dim=-1, keepdim=True)
    v_reg_n = v_reg / torch.norm(v_reg, dim=-1, keepdim=True)
    score = torch.reshape(cls_score, [1, 1, 1, -1]).repeat(1, H, N, 1)
    raw_cls = v_cls_n @ v_cls_n.transpose(-2, -1)
    raw_reg = v_reg_n @ v_reg_n.transpose(-2, -1)
    attn_cls = ((q_cls @ k_cls.transpose(-2, -1)) * scale * score).softmax(dim=-1)
    attn_reg = ((q_reg @ k_reg.transpose(-2, -1)) * scale).softmax(dim=-1)
    attn = (attn_reg + attn_cls) / 2
    x = (attn @ v_cls).transpose(1, 2).reshape(B, N, C)
    x_c = torch.cat([x, v_cls.permute(0, 2, 1, 3).reshape(B, N, C)], dim=-1)
    xr = (attn @ v_reg).transpose(1, 2).reshape(B, N, C)
    x_r = torch.cat([xr, v_reg.permute(0, 2, 1, 3).reshape(B, N, C)], dim=-1)
    r2c, r2o = _round2(attn, raw_cls, raw_reg, H, sim_thresh, conf_sim_thresh)
    return x_c, x_r, r2c, r2o


def msa_yolov(sd, prefix, x_cls, x_reg, cls_score, num_heads=4, sim_thresh=0.75, conf_sim_thresh=0.99,
              reconf=False):
    """MSA_yolov.forward, post_trans.py:1256-1269 (+find_similar_round2 :1238-1254)."""
    tc, to, r2c, r2o = attention_msa(sd, prefix + "msa.", x_cls, x_reg, cls_score, num_heads, 25.0, sim_thresh,
                                     conf_sim_thresh)
    tc = F.linear(tc, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"])[0]
    out_c = F.linear(torch.cat([r2c @ tc, tc], dim=-1), sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])
    out_o = None
    if reconf:
        to = F.linear(to, sd[prefix + "linear1_obj.weight"], sd[prefix + "linear1_obj.bias"])[0]
        out_o = F.linear(torch.cat([r2o @ to, to], dim=-1), sd[prefix + "linear2_obj.weight"],
                         sd[prefix + "linear2_obj.bias"])
    return out_c, out_o


# --------------------------------------------------------------------------- a15
def lap(cost):
    """scipy.optimize.linear_sum_assignment (plain-C restatement).  cost: 2-D tensor/array."""
    c = np.ascontiguousarray(np.asarray(cost, dtype=np.float64))
    nr, nc = c.shape
    k = min(nr, nc)
    a = np.empty(max(k, 1), dtype=np.int64)
    b = np.empty(max(k, 1), dtype=np.int64)
    rc = _lib().oracle_lap(c.ctypes.data, nr, nc, a.ctypes.data, b.ctypes.data)
    if rc != 0:
        raise ValueError("cost matrix is infeasible")
    return a[:k].copy(), b[:k].copy()


# --------------------------------------------------------------------------- a9
def se_module(sd, prefix, reg_feature, edge_feature):
    """SEModule.forward, tscd_matching.py:278-283: per (row, channel) pair gate 2->32->2."""
    q, b, c = reg_feature.shape
    feat = torch.stack([reg_feature, edge_feature], dim=3).view(q * c, 2)
    w = torch.sigmoid(F.linear(F.relu(F.linear(feat, sd[prefix + "fc.0.weight"])), sd[prefix + "fc.2.weight"]))
    w = w.view(q, b, c, 2)
    return reg_feature * w[:, :, :, 0] + edge_feature * w[:, :, :, 1]


def l2norm_attention(sd, prefix, query, key, value, num_heads):
    """PositionMHAttention.forward without boxes (tscd_matching.py:31-60) == MHAttention.forward
    (:159-181): q,k L2-normalised, unscaled softmax, heads merged, no output projection."""
    N, B, C = query.shape
    M = key.shape[0]
    q = F.linear(query, sd[prefix + "q_reg.weight"]).reshape(N, B, num_heads, C // num_heads).permute(1, 2, 0, 3)
    k = F.linear(key, sd[prefix + "k_reg.weight"]).reshape(M, B, num_heads, C // num_heads).permute(1, 2, 0, 3)
    v = F.linear(value, sd[prefix + "v_reg.weight"]).reshape(M, B, num_heads, C // num_heads).permute(1, 2, 0, 3)
    q = q / torch.norm(q, dim=-1, keepdim=True)
    k = k / torch.norm(k, dim=-1, keepdim=True)
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    return (attn @ v).transpose(1, 2).reshape(B, N, C).transpose(0, 1)


def _referring_layer(sd, prefix, identity, tgt, memory, pos, query_pos, edge, query_edge, num_heads):
    """ReferringCrossAttentionLayer.forward_post, tscd_matching.py:566-589."""
    q_in = se_module(sd, prefix + "CA.", tgt, query_edge) + query_pos
    k_in = se_module(sd, prefix + "CA.", memory, edge) + pos
    t2 = l2norm_attention(sd, prefix + "multihead_attn.", q_in, k_in, memory, num_heads)
    x = identity + t2
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])


# --------------------------------------------------------------------------- a8
class CAFMState:
    """The cross-call memory of AwarePositionRegMatcher (tscd_matching.py:708-715)."""

    def __init__(self):
        self.last_output = None      # [n,1,D]  (last_outputs[-1])
        self.last_embeds = None      # only its length is used
        self.last_reg = None
        self.last_cls = None
        self.last_time = None
        self.last_edge = None


def _double_match(ref_a, cur_a, ref_b, cur_b, lap_fn=None, ctx=None):
    """double_match_embds, tscd_matching.py:912-937 (eps 1e-6 in the norms; NaN -> 0).
    lap_fn(cost, ctx), if given, replaces the Hungarian solve (tests: force a measured near-tie the other way)."""
    def cos(r, c):
        r, c = r[:, 0, :], c[:, 0, :]
        r = r / (r.norm(dim=1)[:, None] + 1e-6)
        c = c / (c.norm(dim=1)[:, None] + 1e-6)
        return torch.mm(r, c.transpose(0, 1))
    C = 1 - ((cos(ref_a, cur_a) + cos(ref_b, cur_b)) / 2)
    C = torch.where(torch.isnan(C), torch.full_like(C, 0), C)
    return (lap(C.numpy()) if lap_fn is None else lap_fn(C.numpy(), ctx)), C


def aware_position_reg_matcher(sd, prefix, features, features_reg, features_cls, features_edge, preds_per_frame,
                               time_embedding, resume=False, state=None, num_heads=8, debug=None, lap_fn=None):
    """AwarePositionRegMatcher.forward (CAFM), tscd_matching.py:722-888, decoder_layer_num=1.

    Returns ([Nloc, D] or None, state).  `debug`, if a dict, receives per-frame costs/permutations."""
    state = state if state is not None else CAFMState()
    lp = prefix + "transformer_aware_cross_attention_layers.0."
    if features_edge.shape[-1] != features.shape[-1]:
        features_edge = F.linear(features_edge, sd[prefix + "edge_feature_embedding.weight"],
                                 sd[prefix + "edge_feature_embedding.bias"])
    feats, regs, clss, edges = (t[:, None, :] for t in (features, features_reg, features_cls, features_edge))
    fl, rl, cl, el = [], [], [], []
    s = 0
    for n in preds_per_frame:
        fl.append(feats[s:s + n]); rl.append(regs[s:s + n]); cl.append(clss[s:s + n]); el.append(edges[s:s + n])
        s += n
    te = F.linear(time_embedding, sd[prefix + "absolute_position_embedding.weight"],
                  sd[prefix + "absolute_position_embedding.bias"])[:, None, :]
    outputs, perms = [], []
    prev_perm = None           # permutation of the previous non-empty frame of THIS call (ctx for lap_fn)
    for i, n in enumerate(preds_per_frame):
        if n == 0:
            if i == 0 and resume is False:
                state = CAFMState()
            continue
        ctx = dict(frame=i, prev_perm=prev_perm)
        E, R, Cc, Ed, T = fl[i], rl[i], cl[i], el[i], te[i].unsqueeze(0)
        if (i == 0 and resume is False) or state.last_output is None:      # :779-807
            state.last_embeds, state.last_reg, state.last_cls = E, R, Cc
            (_, col), cost = _double_match(R, R, Cc, Cc, lap_fn, ctx)
            perm = col
            out = _referring_layer(sd, lp, E, E, E, T, T, Ed, Ed, num_heads)
            new_edge = Ed
        else:                                                              # :808-875
            (row, col), cost = _double_match(state.last_reg, R, state.last_cls, Cc, lap_fn, ctx)
            n_prev = len(state.last_embeds)
            if n_prev < n:
                no_match = [j for j in range(n) if j not in col]
                perm = np.append(col, no_match).astype(np.int64)
                last_feat = torch.cat((state.last_output, E[no_match]), dim=0)
                last_edge = torch.cat((state.last_edge, Ed[no_match]), dim=0)
            elif n_prev > n:
                perm = col
                last_feat = state.last_output[row]
                last_edge = state.last_edge[row]
            else:
                perm = col
                last_feat, last_edge = state.last_output, state.last_edge
            state.last_embeds, state.last_reg, state.last_cls = E[perm], R[perm], Cc[perm]
            out = _referring_layer(sd, lp, E[perm], last_feat, E, T, state.last_time, Ed, last_edge, num_heads)
            new_edge = Ed[perm]
        state.last_output, state.last_time, state.last_edge = out, T, new_edge
        if debug is not None:
            debug.setdefault("cost", []).append(cost)
            debug.setdefault("perm", []).append(np.asarray(perm))
        outputs.append(out)
        perms.append(np.asarray(perm))
        prev_perm = np.asarray(perm)
    if not outputs:
        return None, state
    outs = torch.cat([o[np.argsort(p)] for o, p in zip(outputs, perms)], dim=0)   # :881-884
    outs = F.layer_norm(outs, (outs.shape[-1],), sd[prefix + "decoder_norm.weight"], sd[prefix + "decoder_norm.bias"])
    return outs.squeeze(1), state


# --------------------------------------------------------------------------- a10
def task_aligned(sd, prefix, features_reg, features_obj, preds_per_frame, num_heads=8):
    """TaskAligned.forward, tscd_matching.py:1107-1139 (+CrossAttentionLayer.forward_post :421-433)."""
    lp = prefix + "transformer_cross_attention_layers.0."
    reg, obj = features_reg[:, None, :], features_obj[:, None, :]
    outs, s = [], 0
    for n in preds_per_frame:
        if n == 0:
            continue
        tgt, mem = obj[s:s + n], reg[s:s + n]
        t2 = l2norm_attention(sd, lp + "multihead_attn.", tgt, mem, mem, num_heads)
        x = tgt + t2
        outs.append(F.layer_norm(x, (x.shape[-1],), sd[lp + "norm.weight"], sd[lp + "norm.bias"]))
        s += n
    if not outs:
        return None
    o = torch.cat(outs, dim=0)
    return F.layer_norm(o, (o.shape[-1],), sd[prefix + "decoder_norm.weight"], sd[prefix + "decoder_norm.bias"]).squeeze(1)


# --------------------------------------------------------------------------- a11
def decode_reg_preds5(reg_preds, boxes, bbox_xform_clip=math.log(736.0 / 32)):
    """tscd_head.py:914-949: deltas w.r.t. the still-detector box -> xyxy."""
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    dx, dy = reg_preds[:, 0], reg_preds[:, 1]
    dw = torch.clamp(reg_preds[:, 2], max=bbox_xform_clip)
    dh = torch.clamp(reg_preds[:, 3], max=bbox_xform_clip)
    pcx, pcy = dx * w + cx, dy * h + cy
    pw, ph = torch.exp(dw) * w, torch.exp(dh) * h
    out = torch.zeros_like(reg_preds)
    out[:, 0] = pcx - 0.5 * pw
    out[:, 1] = pcy - 0.5 * ph
    out[:, 2] = pcx + 0.5 * pw
    out[:, 3] = pcy + 0.5 * ph
    return out


# --------------------------------------------------------------------------- a12
def postprocess(prediction, num_classes, fc_outputs, conf_output, reg_output, conf_thre=0.001, nms_thre=0.5, debug=None):
    """post_process.py:9-85 (cls_sig=True).  prediction: list of [n,7+C] (or None); returns
    (output, output_ori) lists of [n_det,7] or None.  `debug`, if a list, receives per frame the candidate rows fed to
    batched_nms and the keep lists (None for skipped frames) -- used by the tests to measure decision margins."""
    output = [None] * len(prediction)
    output_ori = [None] * len(prediction)
    for i, det in enumerate(prediction):
        if debug is not None:
            debug.append(None)
        if det is None or det.shape[0] == 0:
            continue
        ori = det.clone()
        det = det.clone()
        cls_conf, cls_pred = torch.max(fc_outputs[i], -1)
        if conf_output is not None:
            det[:, 4] = conf_output[i].sigmoid()
        if reg_output is not None:
            det[:, :4] = reg_output[i]
        det[:, 5] = cls_conf.sigmoid()
        det[:, 6] = cls_pred
        sc = fc_outputs[i].sigmoid()
        r, c = torch.where(sc >= conf_thre)                                   # row-major order
        new = det[r, :7]
        new[:, 6] = c
        new[:, 5] = sc[r, c]
        new = new[new[:, 4] * new[:, 5] >= conf_thre]
        if new.shape[0] == 0:
            continue                                                           # also skips output_ori (:54-55)
        keep = batched_nms(new[:, :4], new[:, 4] * new[:, 5], new[:, 6], nms_thre)
        output[i] = new[keep]
        o7 = ori[:, :7]
        o7 = o7[o7[:, 4] * o7[:, 5] >= conf_thre]
        keep_ori = batched_nms(o7[:, :4], o7[:, 4] * o7[:, 5], o7[:, 6], nms_thre)
        output_ori[i] = o7[keep_ori]
        if debug is not None:
            debug[-1] = dict(cand=new, keep=keep, cand_ori=o7, keep_ori=keep_ori, all_scores=sc, obj=det[:, 4].clone(),
                             box=det[:, :4].clone())
    return output, output_ori


# --------------------------------------------------------------------------- a1
def stage_tscd(sd, decoded, cls_feat, reg_feat, edge_feat, time_embedding, num_classes, lframe, gframe,
               selection="B", select_kwargs=None, nms_thresh=0.5, sim_thresh=0.75, conf_sim_thresh=0.99, heads=4,
               resume=False, state=None, trace=None, lap_fn=None):
    """TSCDHead.forward inference tail, tscd_head.py:374-733 (agg_type='mca', decouple_reg, reconf).

    decoded [F,A,5+C] (after decode_outputs), feature planes [F,A,D].  selection 'B' =
    postprocess_widx (what the shipped TSCD exps run), 'A' = postpro_woclass (gen-1 / north_star).
    Returns (result, result_ori, state)."""
    select_kwargs = dict(select_kwargs or {})
    if selection == "B":
        rows, idxs = select_mode_b(decoded, num_classes, **select_kwargs)
    else:
        rows, idxs = select_mode_a(decoded, num_classes, **select_kwargs)
    ppf = [0 if p is None else int(p.shape[0]) for p in idxs]
    if decoded.shape[0] == 1:
        return rows, rows, state                                             # :429-430
    bank = find_feature_score(cls_feat, reg_feat, edge_feat, idxs, rows, num_classes)
    if bank is None or bank[0].shape[0] == 0:
        return rows, rows, state
    f_cls, f_reg, f_edge, cls_scores, fg_scores, locs, all_scores = bank
    x_cls, x_reg = f_cls.unsqueeze(0), f_reg.unsqueeze(0)
    agg_cls, _ = mca_tscd_g2l_reg(sd, "agg.", x_cls, x_reg, cls_scores, ppf, lframe, heads, sim_thresh, conf_sim_thresh)
    iou_cls, iou_reg = mca_tscd_g2l_reg(sd, "agg_iou.", x_cls, x_reg, cls_scores, ppf, lframe, heads, sim_thresh,
                                        conf_sim_thresh)
    n_loc = sum(ppf[:lframe])
    dbg = {} if trace is not None else None
    matched, state = aware_position_reg_matcher(sd, "local_reg_matcher.", f_reg, iou_reg, iou_cls, f_edge, ppf[:lframe],
                                                time_embedding[:lframe], resume=resume, state=state, debug=dbg, lap_fn=lap_fn)
    if matched is None:
        matched = f_reg[:n_loc]
    matched = F.linear(matched, sd["fc_reg_matcher.weight"], sd["fc_reg_matcher.bias"])          # :507
    obj_ref = task_aligned(sd, "task_aligned.", matched, iou_reg, ppf[:lframe])                   # :508-511
    if obj_ref is None:
        obj_ref = iou_reg
    cls_preds = F.linear(agg_cls, sd["cls_pred.weight"], sd["cls_pred.bias"])                    # :515-520
    obj_preds = F.linear(obj_ref, sd["matcher_obj_pred.weight"], sd["matcher_obj_pred.bias"])
    reg_deltas = F.linear(matched, sd["matcher_reg_pred.weight"], sd["matcher_reg_pred.bias"])
    ori_boxes = torch.cat([rows[i][:, :4] for i in range(lframe) if rows[i] is not None], dim=0)
    reg_preds = decode_reg_preds5(reg_deltas, ori_boxes)                                          # :689-706
    cls_pf, obj_pf, reg_pf, s = [], [], [], 0
    for i in range(lframe):
        n = ppf[i]
        cls_pf.append(cls_preds[s:s + n]); obj_pf.append(obj_preds[s:s + n].squeeze(-1)); reg_pf.append(reg_preds[s:s + n])
        s += n
    if trace is not None:
        trace.update(dict(rows=rows, idxs=idxs, bank=bank, agg_cls=agg_cls, iou_cls=iou_cls, iou_reg=iou_reg,
                          matched=matched, obj_ref=obj_ref, cls_preds=cls_preds, obj_preds=obj_preds,
                          reg_deltas=reg_deltas, reg_preds=reg_preds, cafm=dbg))
    post_dbg = [] if trace is not None else None
    result, result_ori = postprocess([rows[i] for i in range(lframe)], num_classes, cls_pf, obj_pf, reg_pf,
                                     nms_thre=nms_thresh, debug=post_dbg)
    if trace is not None:
        trace["post"] = post_dbg
    return result, result_ori, state


def stage_gen1(sd, decoded, cls_feat, reg_feat, num_classes, pre_k=750, top_k=30, nms_thre=0.75, heads=4,
               sim_thresh=0.75, trace=None):
    """Gen-1 (YOLOV) pair named by the north_star: postpro_woclass -> gather -> MSA_yolov -> cls_pred.
    (yolovp_msa.py:326-376; its own final `postprocess` call is broken upstream, SURVEY finding 3, so
    the stage ends at the refined class logits.)"""
    rows, idxs = select_mode_a(decoded, num_classes, nms_thre=nms_thre, pre_k=pre_k, top_k=top_k)
    bank = find_feature_score(cls_feat, reg_feat, reg_feat, idxs, rows, num_classes)
    f_cls, f_reg, _, cls_scores, fg_scores, _, _ = bank
    out_c, _ = msa_yolov(sd, "trans.", f_cls.unsqueeze(0), f_reg.unsqueeze(0), cls_scores, heads, sim_thresh)
    logits = F.linear(out_c, sd["linear_pred.weight"], sd["linear_pred.bias"])
    if trace is not None:
        trace.update(dict(rows=rows, idxs=idxs, bank=bank, msa=out_c))
    return rows, idxs, logits


# --------------------------------------------------------------------------- synthetic workload
def synth_head_outputs(num_frames, hw, num_classes, dim=256, seed=2024, clustered=False, dtype=torch.float32,
                       obj_mean=-3.0, n_obj=40, n_mem=12):
    """Synthetic boundary tensors at seam S2: pre-decode head outputs [F,A,5+C] with sigmoid applied
    (tscd_head.py:357-359,374-376) and three feature planes [F,A,dim].  obj/cls logits ~ N(-3,2^2),
    dx,dy ~ U(-0.5,1.5), dw,dh ~ N(1.0,0.7^2); `clustered` duplicates n_obj (40) objects per frame over n_mem (12) anchors with jitter;
    `obj_mean` (scalar or per frame) shifts the objectness logits to vary how many anchors pass 0.001."""
    g = torch.Generator().manual_seed(seed)
    A = sum(h * w for h, w in hw)
    obj_mean = torch.as_tensor(obj_mean, dtype=torch.float32).reshape(-1, 1, 1)   # scalar or one per frame
    obj = torch.randn(num_frames, A, 1, generator=g) * 2 + obj_mean
    cls = torch.randn(num_frames, A, num_classes, generator=g) * 2 - 3
    xy = torch.rand(num_frames, A, 2, generator=g) * 2 - 0.5
    wh = torch.randn(num_frames, A, 2, generator=g) * 0.7 + 1.0
    feats = [torch.randn(num_frames, A, dim, generator=g) for _ in range(3)]
    if clustered:
        # objects: members are random anchors whose decoded boxes regress to the same object (+ jitter),
        # share its class and carry near-duplicate features -> pre-NMS really suppresses and the 0.75 /
        # 0.99 cosine masks of the aggregation are non-trivial.
        gx, gy, gs = [], [], []
        for (h, w), s in zip(hw, (8, 16, 32)):
            yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
            gx.append(xx.reshape(-1).float()); gy.append(yy.reshape(-1).float()); gs.append(torch.full((h * w,), float(s)))
        gx, gy, gs = torch.cat(gx), torch.cat(gy), torch.cat(gs)
        size = float(hw[0][1] * 8)
        for f in range(num_frames):
            for o in range(n_obj):
                m = torch.randint(0, A, (n_mem,), generator=g)
                cxy = torch.rand(2, generator=g) * size
                bwh = torch.rand(2, generator=g) * size * 0.25 + 16
                c = int(torch.randint(0, num_classes, (1,), generator=g))
                jit = 0.04 * torch.randn(n_mem, 2, generator=g) * bwh
                xy[f, m, 0] = (cxy[0] + jit[:, 0]) / gs[m] - gx[m]
                xy[f, m, 1] = (cxy[1] + jit[:, 1]) / gs[m] - gy[m]
                wh[f, m] = torch.log(bwh / gs[m][:, None]) + 0.04 * torch.randn(n_mem, 2, generator=g)
                obj[f, m] = 1.0 + torch.randn(n_mem, 1, generator=g)
                cls[f, m, c] = 2.0 + torch.randn(n_mem, generator=g)
                for p in feats:
                    p[f, m] = torch.randn(1, dim, generator=g) + 0.05 * torch.randn(n_mem, dim, generator=g)
    head = torch.cat([xy, wh, torch.sigmoid(obj), torch.sigmoid(cls)], dim=2)
    return head.to(dtype), [p.to(dtype) for p in feats]


def init_stage_weights(num_classes, dim=256, seed=2024, gen1=False):
    """Random-init aggregation-stage weights with the reference's key names and shapes
    (SURVEY App. B).  Distribution: uniform(+-1/sqrt(fan_in)) like nn.Linear's default; LayerNorm = (1,0)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f, bias=True):
        b = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b
        if bias:
            sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * b

    def ln(name, n):
        sd[name + ".weight"] = torch.ones(n) + 0.1 * (torch.rand(n, generator=g) - 0.5)
        sd[name + ".bias"] = 0.1 * (torch.rand(n, generator=g) - 0.5)

    D = dim
    if gen1:
        lin("trans.msa.qkv_cls", 3 * D, D, False); lin("trans.msa.qkv_reg", 3 * D, D, False)
        lin("trans.linear1", 2 * D, 2 * D); lin("trans.linear2", 4 * D, 4 * D)
        lin("linear_pred", num_classes + 1, 4 * D)
        return sd
    for m in ("agg.", "agg_iou."):
        lin(m + "mca.q_cls_local", D, D, False); lin(m + "mca.kv_cls", 2 * D, D, False)
        lin(m + "mca.q_reg_local", D, D, False); lin(m + "mca.kv_reg", 2 * D, D, False)
        lin(m + "mca.linear", 2 * D, 2 * D); lin(m + "mca.linear_reg", 2 * D, 2 * D)
        lin(m + "linear", 4 * D, 3 * D); lin(m + "linear_obj", 4 * D, 3 * D)
    p = "local_reg_matcher."
    for layer in ("transformer_self_attention_layers.0.self_attn.", "transformer_aware_cross_attention_layers.0.multihead_attn."):
        for n in ("q_reg", "k_reg", "v_reg"):
            lin(p + layer + n, D, D, False)
        sd[p + layer + "position_embedding.weight"] = torch.randn(8, 64, 1, 1, generator=g) * 0.1
        sd[p + layer + "position_embedding.bias"] = torch.zeros(8)
    for layer in ("transformer_self_attention_layers.0.", "transformer_aware_cross_attention_layers.0."):
        ln(p + layer + "norm", D)
        lin(p + layer + "CA.fc.0", 32, 2, False); lin(p + layer + "CA.fc.2", 2, 32, False)
    lin(p + "transformer_ffn_layers.0.linear1", D, D); lin(p + "transformer_ffn_layers.0.linear2", D, D)
    ln(p + "transformer_ffn_layers.0.norm", D)
    lin(p + "absolute_position_embedding", D, 256); lin(p + "edge_feature_embedding", D, D // 4)
    ln(p + "decoder_norm", D)
    lin("fc_reg_matcher", 4 * D, D)
    t = "task_aligned.transformer_cross_attention_layers.0."
    for n in ("q_reg", "k_reg", "v_reg"):
        lin(t + "multihead_attn." + n, 4 * D, 4 * D, False)
    ln(t + "norm", 4 * D); ln("task_aligned.decoder_norm", 4 * D)
    lin("cls_pred", num_classes, 4 * D); lin("matcher_obj_pred", 1, 4 * D); lin("matcher_reg_pred", 4, 4 * D)
    return sd
