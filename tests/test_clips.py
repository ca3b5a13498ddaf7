"""CPU tests of the host logic before the stage: clip construction / frame sampling / time embedding / CAFM-memory
scheduling of tscd_b200.clips against goldens produced by running the reference's own code (tools/make_goldens_host.py)."""
import json
import os
import random

import torch

from conftest import GOLDEN


def _cases():
    return json.load(open(os.path.join(GOLDEN, "clips.json")))


def test_dataset_clips_match_reference_photo_to_sequence():
    from tscd_b200 import clips
    n = 0
    for c in _cases():
        if c["fn"] == "demo":
            continue
        if c["fn"] == "ovis":
            videos = [[f"v{sid}/img_{k:07d}.jpg" for k in range(ln)] for sid, ln in enumerate(c["lengths"])]
        else:
            videos = [[f"v{sid}/{k:06d}.JPEG" for k in range(ln)] for sid, ln in enumerate(c["lengths"])]
        random.seed(c["seed"])
        got = clips.dataset_clips(videos, c["lframe"], c["gframe"], mode=c["mode"], dataset=c["fn"], formal=c.get("formal", False))
        assert got == c["clips"], (c["fn"], c["mode"], c["lframe"], c["gframe"])
        n += 1
    assert n == 9


def test_demo_clips_match_reference_loop():
    from tscd_b200 import clips
    n = 0
    for c in _cases():
        if c["fn"] != "demo":
            continue
        random.seed(c["seed"])
        res, seq = clips.demo_clips(c["frame_len"], c["lframe"], c["gframe"])
        assert res == c["clips"], c
        assert seq == c["path_sequence"], c
        n += 1
    assert n == 6


def test_frame_number_time_embedding_resume():
    from tscd_b200 import clips, weights
    import oracle
    assert clips.frame_number("data/val/v3/000123.JPEG", "vid") == 123
    assert clips.frame_number("data/ovis/valid/abc/img_0000042.jpg", "ovis") == 42
    te = clips.time_embedding([0, 1, 2, 17, 255])
    assert te.shape == (5, 256)
    assert torch.equal(te, oracle.timing_signal_1d(torch.tensor([0, 1, 2, 17, 255]), 256))      # oracle pinned by stage_tscd.npz
    assert torch.equal(te, weights.timing_signal_1d(torch.tensor([0, 1, 2, 17, 255]), 256))
    assert clips.resume_flag(8, 0) is False and clips.resume_flag(8, 8) is True and clips.resume_flag(0, 40) is False


def test_clip_scheduler_keeps_videos_in_order_on_one_rank():
    from tscd_b200 import clips
    random.seed(3)
    lengths = [100, 37, 64, 20, 9, 80, 55]
    per_video = []
    for n in lengths:
        res, seq = clips.demo_clips(n, 8, 24 if n > 32 else max(1, n - 8))
        per_video.append(list(zip(res, seq)))
    for world, slots in ((1, 1), (2, 3), (8, 2)):
        sch = clips.ClipScheduler(per_video, world=world, slots=slots)
        seen = {}
        for r in range(world):
            last_in_slot = {}
            for batch in sch.batches(r):
                assert 1 <= len(batch) <= slots
                assert len({c.slot for c in batch}) == len(batch)
                for c in batch:
                    seen.setdefault(c.video, []).append((r, c.clip))
                    prev = last_in_slot.get(c.slot)
                    if c.resume:                       # continues exactly the previous clip of the same video in the same slot
                        assert prev == (c.video, c.clip - 1)
                    else:
                        assert c.clip == 0
                    last_in_slot[c.slot] = (c.video, c.clip)
                    assert c.frames == per_video[c.video][c.clip][0] and c.frame_numbers == per_video[c.video][c.clip][1]
        assert sorted(seen) == list(range(len(lengths)))
        for v, lst in seen.items():
            assert len({r for r, _ in lst}) == 1                       # one rank per video
            assert [k for _, k in lst] == list(range(len(per_video[v])))   # every clip once, in order
        loads = [sum(len(per_video[v]) for v in sch.assignment[r]) for r in range(world)]
        assert max(loads) - min(loads) <= max(len(p) for p in per_video)
