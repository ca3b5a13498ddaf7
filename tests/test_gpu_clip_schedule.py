"""SURVEY 8f-4 on the device: clips of several videos batched by tscd_b200.clips.ClipScheduler (one CAFMState slot per concurrent
video, resume = 0 on a video's first clip, slots selected / written back when fewer streams than slots remain) must give exactly
the detections of processing every video alone, clip after clip."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def test_scheduled_batches_equal_sequential_videos():
    from tscd_b200 import clips, ops, selection, stage
    F, Lf, C, D = 6, 2, 5, 256
    hw = [(16, 16), (8, 8), (4, 4)]
    dt = torch.float16
    sd = oracle.init_stage_weights(C, dim=D, seed=3)
    cfg = stage.StageConfig(num_classes=C, selection=selection.SelectionConfig(mode="B", minimal_limit=12, maximal_limit=40, use_pre_nms=False), dtype=dt)
    st = stage.AggregationStage(cfg, sd)
    an = ops.AnchorSpec(hw)
    kmax = cfg.selection.max_keep(an.num_anchors)
    n_clips = [3, 2, 1, 2]
    data = {}
    for v, n in enumerate(n_clips):
        for c in range(n):
            h, f = oracle.synth_head_outputs(F, hw, C, dim=D, seed=1000 + 10 * v + c, clustered=True, obj_mean=-7.0)
            data[(v, c)] = (oracle.decode_outputs(h, hw, [8, 16, 32]).cuda(), [p.to(dt).cuda().contiguous() for p in f])

    def run(batch, state, resume):
        head = ops.HeadViews.from_fused(torch.cat([data[k][0] for k in batch], 0), an, apply_sigmoid=False, apply_decode=False)
        feats = [torch.cat([data[k][1][i] for k in batch], 0).contiguous() for i in range(3)]
        views = tuple(ops.view_rowmajor(t, an) for t in feats)
        te = torch.cat([oracle.timing_signal_1d(torch.arange(c * Lf, (c + 1) * Lf), 256) for (_, c) in batch], 0)
        out = st.forward(head, views, dt, te, len(batch), F, Lf, state=state, resume=torch.tensor(resume, dtype=torch.int32, device="cuda"))
        torch.cuda.synchronize()
        res, ori = st.to_lists(out, len(batch), Lf)
        return [(res[i * Lf:(i + 1) * Lf], ori[i * Lf:(i + 1) * Lf]) for i in range(len(batch))]

    want = {}
    for v, n in enumerate(n_clips):                        # every video alone, its own memory
        state = stage.CAFMState(1, kmax, D)
        for c in range(n):
            want[(v, c)] = run([(v, c)], state, [0 if c == 0 else 1])[0]

    per_video = [[([v, c], list(range(c * Lf, (c + 1) * Lf))) for c in range(n)] for v, n in enumerate(n_clips)]
    sch = clips.ClipScheduler(per_video, world=1, slots=2)
    state = stage.CAFMState(2, kmax, D)
    seen = 0
    for batch in sch.batches(0):
        keys = [(c.video, c.clip) for c in batch]
        slots = [c.slot for c in batch]
        sub = state if slots == [0, 1] else state.select(slots)
        got = run(keys, sub, [c.resume for c in batch])
        if sub is not state:
            state.update_from(sub, slots)
        for k, (r, o) in zip(keys, got):
            for a, b in zip(r + o, want[k][0] + want[k][1]):
                assert (a is None) == (b is None), k
                if a is not None:
                    assert torch.equal(a, b), k
            seen += 1
    assert seen == sum(n_clips)
