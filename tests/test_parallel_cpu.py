"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: clip sharding, frame ownership of the long clip, the single
all-gather of the packed bank buffers and the one-message CAFM state hand-over between ranks."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_clips_contiguous_and_complete():
    from tscd_b200.parallel import shard_clips
    for n in (1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_clips(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


class _FlatState:
    """Stand-in with CAFMState's hand-over interface (one flat buffer) for the CPU test."""

    def __init__(self, fill):
        self.flat = torch.full((1000,), float(fill))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tscd_b200 import parallel
        # ONE all-gather of a packed byte buffer per rank (the long-clip exchange; the CUDA pack / unpack kernels around it are
        # covered by tests/test_gpu_long_clip.py)
        send = torch.full((4096,), rank + 1, dtype=torch.uint8)
        recv = parallel.all_gather_bytes(send)
        assert recv.shape == (world * 4096,)
        for r in range(world):
            assert bool((recv[r * 4096:(r + 1) * 4096] == r + 1).all())
        # frame ownership: consecutive blocks, complete, in rank order
        loc, glob = [], []
        for r in range(world):
            l_, g_ = parallel.frame_plan(8, 24, r, world)
            loc += l_; glob += g_
        assert loc == list(range(8)) and glob == list(range(8, 32))
        # CAFM state hand-over rank 0 -> rank 1: one message
        st = _FlatState(3.0 if rank == 0 else 0.0)
        if rank == 0:
            parallel.send_state(st, 1)
        else:
            parallel.recv_state(st, 0)
            assert bool((st.flat == 3.0).all())
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_long_clip_exchange_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_cafm_state_is_one_flat_buffer():
    """Every field of CAFMState is a view of `flat` (so the rank-to-rank hand-over is a single message)."""
    from tscd_b200.stage import CAFMState
    st = CAFMState(2, 5, 8, device="cpu")
    st.flat.fill_(1.0)
    for f in ("out", "edge", "reg", "cls", "nreg", "ncls", "time"):
        t = getattr(st, f)
        assert bool((t == 1.0).all()) and t.untyped_storage().data_ptr() == st.flat.untyped_storage().data_ptr()
    st.n.fill_(7)
    assert st.flat[:2].view(torch.int32).tolist() == [7, 7] and st.out.shape == (2, 5, 8) and st.reg.shape == (2, 5, 32)
