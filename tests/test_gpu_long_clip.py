"""GPU tests of the long-clip mode (BASELINE configs[3]) with the ranks EMULATED on one GPU: the pack / unpack kernels of the
bank exchange are exact, and a clip processed as W frame shards (own local frames + all-gathered global rows, CAFM memory
handed shard to shard) gives the detections of the same clip processed in one piece.  The real NCCL path is
tools/check_long_clip.py / bench.py --mode long-clip under torchrun."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _clip(F, hw, C, seed):
    h, f = oracle.synth_head_outputs(F, hw, C, dim=256, seed=seed, clustered=True, obj_mean=[-7.5 + 0.5 * (i % 5) for i in range(F)])
    return oracle.decode_outputs(h, hw, [8, 16, 32]), [p.half() for p in f]


def _select(frames, decoded, feats, an, cfg, kmax):
    from tscd_b200 import ops, selection
    h = ops.HeadViews.from_fused(decoded[frames].cuda(), an, apply_sigmoid=False, apply_decode=False)
    fv = [f[frames].cuda().contiguous() for f in feats]
    views = tuple(ops.view_rowmajor(f, an) for f in fv)
    rows_cap = ((len(frames) * kmax + 127) // 128) * 128 + 128
    return selection.select_and_gather(h, views, torch.float16, 256, cfg, bank_dtype=torch.float16, bank_rows=rows_cap), (h, fv)


@pytest.mark.parametrize("mode,world", [("B", 4), ("A", 8)])
def test_long_clip_shards_match_single_pass(mode, world):
    from tscd_b200 import ops, parallel, selection, stage, weights
    C, hw = 5, [(16, 16), (8, 8), (4, 4)]
    L, G = 2 * world, 3 * world
    F = L + G
    an = ops.AnchorSpec(hw)
    decoded, feats = _clip(F, hw, C, seed=90 + world)
    sel_cfg = (selection.SelectionConfig(mode="B", minimal_limit=8, maximal_limit=24, use_pre_nms=False) if mode == "B"
               else selection.SelectionConfig(mode="A", pre_k=120, top_k=16))
    cfg = stage.StageConfig(num_classes=C, selection=sel_cfg)
    st = stage.AggregationStage(cfg, weights.random_state_dict(C, 256, seed=5))
    kmax = sel_cfg.max_keep(an.num_anchors)
    te = weights.timing_signal_1d(torch.arange(L), 256)
    # ---- reference: the whole clip in one piece ----
    sel_full, keep0 = _select(list(range(F)), decoded, feats, an, sel_cfg, kmax)
    out_full = st.forward_from_bank(sel_full, 1, F, L, kmax, te, state=stage.CAFMState(1, kmax, 256))
    torch.cuda.synchronize()
    full, full_ori = st.to_lists(out_full, 1, L)
    # ---- emulated ranks ----
    sels, sends, keep = [], [], []
    for r in range(world):
        loc, glob = parallel.frame_plan(L, G, r, world)
        s_r, k_r = _select(loc + glob, decoded, feats, an, sel_cfg, kmax)
        sels.append(s_r); keep.append(k_r)
        sends.append(parallel.pack_global_bank(s_r, L // world, G // world, kmax, torch.float16))
    recv = torch.cat(sends)                       # what ONE all-gather delivers to every rank
    state = stage.CAFMState(1, kmax, 256)          # handed rank -> rank (one flat buffer)
    sharded, sharded_ori = [], []
    full_cnt = sel_full["sel_count"].cpu().tolist()
    full_off = sel_full["row_off"].cpu().tolist()
    for r in range(world):
        virt, F_virt = parallel.unpack_virtual_clip(sels[r], recv, world, L // world, G // world, kmax, torch.float16)
        torch.cuda.synchronize()
        Lr = L // world
        # exactness of the exchange: counts, offsets and rows of the virtual clip == own local frames | all global frames
        want_cnt = full_cnt[r * Lr:(r + 1) * Lr] + full_cnt[L:]
        assert virt["sel_count"].cpu().tolist() == want_cnt
        off = virt["row_off"].cpu().tolist()
        assert off == [0] + torch.tensor(want_cnt).cumsum(0).tolist()
        for key in ("bank_cls", "bank_reg", "bank_score"):
            want = torch.cat([sel_full[key][full_off[r * Lr]:full_off[(r + 1) * Lr]], sel_full[key][full_off[L]:full_off[F]]])
            assert torch.equal(virt[key][:off[-1]], want), (key, r)
        n_loc = off[Lr]
        assert torch.equal(virt["bank_edge"][:n_loc], sel_full["bank_edge"][full_off[r * Lr]:full_off[(r + 1) * Lr]])
        resume = torch.tensor([1 if r > 0 else 0], dtype=torch.int32).cuda()
        out = st.forward_from_bank(virt, 1, F_virt, Lr, kmax, te[r * Lr:(r + 1) * Lr], state=state, resume=resume)
        torch.cuda.synchronize()
        a, b = st.to_lists(out, 1, Lr)
        sharded += a; sharded_ori += b
    tot = match = 0
    for got, want in zip(sharded + sharded_ori, full + full_ori):
        assert (got is None) == (want is None)
        if got is None:
            continue
        g, w = got.cpu(), want.cpu()
        assert len(g) == len(w), "detection counts per frame must agree exactly"
        tot += len(w)
        used = set()
        for i in range(len(w)):
            for j in range(len(g)):
                if j not in used and g[j, 6] == w[i, 6] and float((g[j, :4] - w[i, :4]).abs().max()) <= 0.3 \
                        and torch.allclose(g[j, 4:6], w[i, 4:6], rtol=1e-2, atol=1e-5):
                    used.add(j); match += 1
                    break
    print(f"long clip, {world} emulated ranks, mode {mode}: {match}/{tot} detections identical (boxes 0.3 px, scores 1e-2)")
    assert tot > 50 and match / tot >= 0.999
