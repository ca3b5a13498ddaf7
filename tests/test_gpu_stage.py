"""GPU parity of the whole aggregation stage (K1..final NMS) through the C-ABI vs the oracle, on a batch of
clips with ragged proposal counts, including the CAFM recurrence with resume across two consecutive calls.

Bars: selection ids exact; Hungarian permutations exact wherever the optimum is not a near-tie; float tensors max-normalised error <= 1e-2 (fp16
tensor-core operands, fp32 accumulation vs the fp32 oracle fed the same 16-bit-rounded inputs/weights);
final detections identical as (frame, class) multisets with boxes/scores within tolerance."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def _round_sd(sd, dtype):
    return {k: (v.to(dtype).float() if v.dim() == 2 and "CA.fc" not in k else v.clone()) for k, v in sd.items()}


def _run_case(mode, B, F, Lf, hw, C, sel_kw, o_sel_kw, seeds, calls=1, dtype=torch.float16):
    from tscd_b200 import ops, selection, stage
    D = 256
    sd = oracle.init_stage_weights(C, dim=D, seed=17)
    # make the prediction heads decisive so the 0.001 filters / final NMS do real work
    sd["cls_pred.weight"] = sd["cls_pred.weight"] * 30.0
    sd["cls_pred.bias"] = sd["cls_pred.bias"] - 4.0
    sd["matcher_obj_pred.weight"] = sd["matcher_obj_pred.weight"] * 5.0
    sd["matcher_obj_pred.bias"] = sd["matcher_obj_pred.bias"] - 1.0
    sd16 = _round_sd(sd, dtype)
    cfg = stage.StageConfig(num_classes=C, selection=selection.SelectionConfig(mode=mode, **sel_kw), dtype=dtype)
    st = stage.AggregationStage(cfg, sd)
    an = ops.AnchorSpec(hw)
    state = None
    o_states = [None] * B
    report = []
    for call in range(calls):
        heads, planes = [], []
        for b in range(B):
            h, f = oracle.synth_head_outputs(F, hw, C, dim=D, seed=seeds[call] + b, clustered=True,
                                             obj_mean=[-7.5 + 0.6 * ((b + i) % 4) for i in range(F)])
            heads.append(oracle.decode_outputs(h, hw, [8, 16, 32]))
            planes.append([p.to(dtype).float() for p in f])
        decoded = torch.cat(heads, 0)
        feats = [torch.cat([planes[b][k] for b in range(B)], 0) for k in range(3)]
        te = torch.cat([oracle.timing_signal_1d(torch.arange(call * Lf, call * Lf + Lf), 256) for _ in range(B)], 0)
        head = ops.HeadViews.from_fused(decoded.cuda(), an, apply_sigmoid=False, apply_decode=False)
        dev_feats = [f.to(dtype).cuda().contiguous() for f in feats]
        views = tuple(ops.view_rowmajor(f, an) for f in dev_feats)
        trace = {}
        resume = torch.full((B,), int(call > 0), dtype=torch.int32).cuda()
        out = st.forward(head, views, dtype, te, B, F, Lf, state=state, resume=resume, trace=trace)
        state = out["state"]
        torch.cuda.synchronize()
        res, res_ori = st.to_lists(out, B, Lf)
        lrow = out["layout"].lrow_off.cpu().tolist()
        perm = trace["perm"].cpu().numpy()
        te16 = te.to(dtype).float()
        for b in range(B):
            otr = {}
            o_res, o_ori, o_states[b] = oracle.stage_tscd(
                sd16, heads[b], planes[b][0], planes[b][1], planes[b][2], te16[b * Lf:(b + 1) * Lf], C, Lf, F - Lf,
                selection=mode, select_kwargs=o_sel_kw, nms_thresh=0.5, resume=(call > 0), state=o_states[b], trace=otr)
            # selection: exact
            cnt = out["sel"]["sel_count"].cpu().tolist()
            for f in range(F):
                n = cnt[b * F + f]
                want = otr["idxs"][f]
                assert out["sel"]["sel_idx"][b * F + f, :n].cpu().tolist() == (want.tolist() if want is not None else [])
            l0, l1 = lrow[b * Lf], lrow[(b + 1) * Lf]
            errs = dict(
                agg_cls=_rel(trace["agg_cls"][l0:l1], otr["agg_cls"]),
                iou_cls=_rel(trace["iou_cls"][l0:l1], otr["iou_cls"]),
                iou_reg=_rel(trace["iou_reg"][l0:l1], otr["iou_reg"]),
                matched=_rel(trace["matched"][l0:l1], otr["matched"]),
                obj_ref=_rel(trace["obj_ref"][l0:l1], otr["obj_ref"]),
                cls_logits=_rel(trace["cls_logits"][l0:l1, :C], otr["cls_preds"]),
                obj_logits=_rel(trace["obj_logits"][l0:l1, :1], otr["obj_preds"]),
                reg_deltas=_rel(trace["reg_deltas"][l0:l1, :4], otr["reg_deltas"]),
            )
            # Hungarian permutations: exact
            o_perm = np.concatenate(otr["cafm"]["perm"]) if otr["cafm"].get("perm") else np.zeros(0)
            perm_ok = np.array_equal(perm[l0:l1], o_perm)
            # detections
            tot = match = 0
            for f in range(Lf):
                for got, want in ((res[b * Lf + f], o_res[f]), (res_ori[b * Lf + f], o_ori[f])):
                    if want is None or got is None:
                        assert want is None and got is None
                        continue
                    g, w_ = got.cpu(), want
                    tot += max(len(g), len(w_))
                    used = set()
                    for i in range(len(w_)):
                        cand = [j for j in range(len(g)) if j not in used and g[j, 6] == w_[i, 6]
                                and torch.allclose(g[j, :6], w_[i, :6], rtol=2e-2, atol=0.75)]
                        if cand:
                            used.add(cand[0]); match += 1
            report.append((call, b, errs, perm_ok, match, tot))
    return report


def _check(report):
    for call, b, errs, perm_ok, match, tot in report:
        print(f"call {call} clip {b}: perm_ok={perm_ok} dets {match}/{tot} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    # The matching costs come from fp16-operand GEMMs here, so a near-tie assignment may legitimately flip
    # (exact assignment parity on exact costs is tests/test_gpu_cafm.py); require it for most clips.
    assert sum(1 for r in report if r[3]) >= 0.6 * len(report)
    for call, b, errs, perm_ok, match, tot in report:
        for k, v in errs.items():
            assert v < 1e-2, f"{k} rel err {v} (call {call}, clip {b})"
        assert tot == 0 or match / tot >= 0.98, f"detections {match}/{tot}"


def test_stage_mode_b_ragged_with_resume():
    """What the shipped TSCD-L exps run: postprocess_widx limits, no pre-NMS; two consecutive clip batches."""
    rep = _run_case("B", B=3, F=6, Lf=3, hw=[(16, 16), (8, 8), (4, 4)], C=5,
                    sel_kw=dict(minimal_limit=12, maximal_limit=40, use_pre_nms=False),
                    o_sel_kw=dict(nms_thre=0.75, minimal_limit=12, maximal_limit=40, use_pre_nms=False),
                    seeds=[100, 200], calls=2)
    _check(rep)


def test_stage_mode_a_topk_nms():
    """BASELINE config 2 selection: top-750 objectness -> class-aware NMS(0.75) -> first 30, then TSCD MCA/CAFM."""
    rep = _run_case("A", B=2, F=8, Lf=2, hw=[(40, 40), (20, 20), (10, 10)], C=25,
                    sel_kw=dict(pre_k=750, top_k=30, nms_thresh=0.75),
                    o_sel_kw=dict(nms_thre=0.75, pre_k=750, top_k=30), seeds=[7], calls=1)
    _check(rep)


def test_stage_full_size_baseline_config():
    """BASELINE.json configs[1] at full size: one 32-frame clip (8 local + 24 global) at 576x576 (6804 anchors), 25
    classes, top-750 -> NMS 0.75 -> 30 proposals/frame, through every kernel of the stage (split K1, top-K NMS prefix
    path, tcgen05 attention, smem/mma.sync CAFM chain, register LSAP, mma.sync TaskAligned, per-class final NMS)."""
    rep = _run_case("A", B=1, F=32, Lf=8, hw=[(72, 72), (36, 36), (18, 18)], C=25,
                    sel_kw=dict(pre_k=750, top_k=30, nms_thresh=0.75),
                    o_sel_kw=dict(nms_thre=0.75, pre_k=750, top_k=30), seeds=[2024], calls=1)
    _check(rep)


def test_stage_bf16_operands():
    """StageConfig.dtype = bfloat16 (tensor-core operands bf16 instead of fp16): every kernel has a bf16 instantiation.
    bf16 keeps 8 mantissa bits, so the float tolerance is 8x looser and near-threshold decisions (0.75 / 0.99 cosine masks,
    Hungarian near-ties, 0.001 score filters) may flip: only a loose detection agreement is required."""
    rep = _run_case("A", B=2, F=8, Lf=2, hw=[(40, 40), (20, 20), (10, 10)], C=25,
                    sel_kw=dict(pre_k=750, top_k=30, nms_thresh=0.75),
                    o_sel_kw=dict(nms_thre=0.75, pre_k=750, top_k=30), seeds=[7], calls=1, dtype=torch.bfloat16)
    for call, b, errs, perm_ok, match, tot in rep:
        print(f"bf16 call {call} clip {b}: perm_ok={perm_ok} dets {match}/{tot} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
        for k, v in errs.items():
            assert v < 5e-2, f"{k} rel err {v}"
        assert tot == 0 or match / tot >= 0.85
