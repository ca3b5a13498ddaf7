"""Multi-GPU modes of the aggregation stage.

* clip-parallel (configs 3, 5): clips are independent units -> `shard_clips` gives every rank a contiguous block; there
  is NO data-path collective (bench.py --gpus N).
* one long clip sharded by frame (config 4): every rank selects / gathers ITS frames (K1-K3 are per frame), then the
  packed bank rows of the GLOBAL frames are exchanged with ONE all-gather (counts first: the rows are ragged) so that
  each rank holds [its own local rows | every global row].  Queries stay sharded, so the attention needs no reduction.
  The CAFM recurrence is a chain over consecutive local frames: rank r receives the CAFM memory from rank r-1 before its
  chain and sends it on afterwards (two tiny point-to-point messages per rank instead of gathering all local rows).

The exchange helpers use torch ops only (device-agnostic, no host sync), so they are exercised with gloo on CPU in
tests/test_parallel_cpu.py; on GPUs the backend is NCCL over NVLink/NVSwitch.
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist


def shard_clips(num_clips: int, rank: int, world: int):
    """Contiguous block of clips for `rank` (consecutive clips of one video stay together and in order)."""
    per = (num_clips + world - 1) // world
    lo = min(rank * per, num_clips)
    return lo, min(lo + per, num_clips)


def exchange_global_bank(sel: Dict[str, torch.Tensor], n_local_frames: int, n_global_frames: int, kmax: int,
                         group=None):
    """All-gather the packed bank rows of this rank's global frames.

    sel: output of selection.select_and_gather for this rank's frames ordered [local frames | global frames]
         (bank_* packed in that order, sel_count [Lr+Gr], row_off [Lr+Gr+1]).
    Returns the `sel` dict of the VIRTUAL clip [own local frames | all ranks' global frames (rank-major)]:
    bank_cls/reg/edge/score, sel_count [Lr + W*Gr], row_off, sel_rows (own frames' rows; only local ones are read).
    Everything stays on the device; the number of valid rows is never read by the host."""
    world = dist.get_world_size(group)
    Lr, Gr = n_local_frames, n_global_frames
    dev = sel["bank_cls"].device
    cnt = sel["sel_count"].to(torch.int64)
    row_off = sel["row_off"].to(torch.int64)
    n_loc = row_off[Lr]                                   # device scalar: own local rows
    cap_g = Gr * kmax                                     # padded rows per rank in the exchange
    ar = torch.arange(cap_g, device=dev)
    src = torch.clamp(n_loc + ar, max=sel["bank_cls"].shape[0] - 1)

    def gather_rows(t):
        send = t.index_select(0, src).contiguous()        # own global rows, packed at the front of a padded block
        recv = torch.empty((world * cap_g,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        return recv

    counts_all = torch.empty(world * Gr, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_all, cnt[Lr:Lr + Gr].contiguous(), group=group)
    n_g = counts_all.view(world, Gr).sum(1)               # valid global rows per rank
    idx = torch.arange(world * cap_g, device=dev)
    valid = (idx % cap_g) < n_g[idx // cap_g]
    order = torch.argsort((~valid).to(torch.int8), stable=True)   # valid rows first, rank-major, original order

    F_virt = Lr + world * Gr
    rows_cap = ((F_virt * kmax + 127) // 128) * 128 + 128
    dest = torch.clamp(n_loc + torch.arange(world * cap_g, device=dev), max=rows_cap - 1)
    out = {}
    for k in ("bank_cls", "bank_reg", "bank_edge", "bank_score"):
        t = sel[k]
        g = gather_rows(t).index_select(0, order)
        v = torch.zeros((rows_cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        n_copy = min(Lr * kmax, t.shape[0], rows_cap)
        v[:n_copy] = t[:n_copy]                           # own local rows sit at the front (plus junk that is overwritten / unused)
        v.index_copy_(0, dest, g)                         # global rows start right after the local ones
        out[k] = v
    counts_virt = torch.cat([cnt[:Lr], counts_all]).to(torch.int32)
    ro = torch.zeros(F_virt + 1, dtype=torch.int32, device=dev)
    ro[1:] = torch.cumsum(counts_virt, 0)
    out["sel_count"], out["row_off"] = counts_virt, ro
    out["sel_rows"], out["sel_idx"] = sel["sel_rows"], sel.get("sel_idx")
    return out, F_virt


_STATE_FIELDS = ("n", "out", "edge", "reg", "cls", "nreg", "ncls", "time")


def send_state(state, dst: int, group=None):
    for f in _STATE_FIELDS:
        dist.send(getattr(state, f).contiguous(), dst, group=group)


def recv_state(state, src: int, group=None):
    for f in _STATE_FIELDS:
        t = getattr(state, f)
        buf = torch.empty_like(t)
        dist.recv(buf, src, group=group)
        t.copy_(buf)


def long_clip_forward(stage, sel, n_local_frames: int, n_global_frames: int, kmax: int, time_embedding_local,
                      state, resume_first: bool = False, group=None, trace: Optional[dict] = None):
    """Frame-sharded forward of ONE long clip.  `sel` = this rank's selection (frames ordered [local | global]);
    ranks hold consecutive blocks of the clip's local frames in rank order.  Returns the stage output for this rank's
    local frames (use AggregationStage.to_lists(out, 1, n_local_frames))."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    virt, F_virt = exchange_global_bank(sel, n_local_frames, n_global_frames, kmax, group)
    dev = virt["bank_cls"].device
    resume = torch.tensor([1 if (rank > 0 or resume_first) else 0], dtype=torch.int32, device=dev)

    def before(st):
        if rank > 0:
            recv_state(st, rank - 1, group)

    def after(st):
        if rank < world - 1:
            send_state(st, rank + 1, group)

    return stage.forward_from_bank(virt, 1, F_virt, n_local_frames, kmax, time_embedding_local, state=state,
                                   resume=resume, trace=trace, before_cafm=before, after_cafm=after)
