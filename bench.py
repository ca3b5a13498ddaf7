#!/usr/bin/env python
"""bench.py -- clip-frames/s of the TSCD aggregation stage (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips B]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): TSCD-L OVIS, 25 classes, 32-frame clips (8 local + 24 global), 576x576
(6804 anchors/frame), pre-NMS top-750 by objectness -> class-aware NMS 0.75 -> 30 proposals/frame, then the TSCD
aggregation (agg + agg_iou MCA, CAFM, TaskAligned, prediction heads, final per-class NMS 0.5).  Synthetic
random-init: head logits/features at seam S1 (raw per-level conv outputs: NCHW fp16 logits, channels_last fp16
feature planes), weights uniform(+-1/sqrt(fan_in)).  A step = one pass of the stage over `--clips` clips per GPU.

value : whole-job clip-frames/s with the step's inputs resident in HBM (inputs per step >> L2: 348 MB per clip).
e2e   : same metric through AggregationStage.forward with HOST (pinned) inputs: H2D of all boundary tensors and
        D2H of the detections inside the timed region.
--impl reference : the reference's CPU path for the same stage (the oracle port of its PyTorch code; the
        reference itself cannot travel to the GPU box), all host threads, on a bounded sample of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F, LF, C, D = 32, 8, 25, 256
HW = [(72, 72), (36, 36), (18, 18)]
PRE_K, TOP_K = 750, 30
WORKLOAD = ("TSCD-L OVIS 25cls, 32-frame clip (8 local + 24 global) @576x576 (6804 anchors), pre-NMS top-750 -> "
            "NMS0.75 -> 30 proposals/frame, agg+agg_iou MCA + CAFM + TaskAligned + final NMS0.5")


# Memory format of the head logits at the seam.  The drop-in head (tscd_b200/head.py) runs the reference's conv towers in
# channels_last, so reg/obj/cls_preds emit channels_last tensors: an anchor's C class logits are one contiguous row and
# mode A reads them for the ~750 survivors only.  --nchw-logits benchmarks PyTorch's default NCHW planes instead (class
# planes streamed by classmax_kernel).
LOGITS_CHANNELS_LAST = True


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------- synthetic inputs
def synth_s1(B, device, seed, pin=False):
    """Seam S1 tensors for B clips: per level reg [BF,4,H,W], obj [BF,1,H,W], cls [BF,C,H,W] (NCHW fp16 logits) and
    three feature planes [BF,256,H,W] (channels_last fp16)."""
    g = torch.Generator(device=device).manual_seed(seed)
    n = B * F
    out = dict(reg=[], obj=[], cls=[], f_cls=[], f_reg=[], f_edge=[])
    for (h, w) in HW:
        xy = torch.rand(n, 2, h, w, generator=g, device=device) * 2 - 0.5
        wh = torch.randn(n, 2, h, w, generator=g, device=device) * 0.7 + 1.0
        cl = torch.channels_last if LOGITS_CHANNELS_LAST else torch.contiguous_format
        out["reg"].append(torch.cat([xy, wh], 1).half().contiguous(memory_format=cl))
        out["obj"].append((torch.randn(n, 1, h, w, generator=g, device=device) * 2 - 3).half().contiguous(memory_format=cl))
        out["cls"].append((torch.randn(n, C, h, w, generator=g, device=device) * 2 - 3).half().contiguous(memory_format=cl))
        for k in ("f_cls", "f_reg", "f_edge"):
            t = torch.empty(n, D, h, w, dtype=torch.float16, device=device).contiguous(memory_format=torch.channels_last)
            for i in range(0, n, 64):       # chunked: bounds the fp32 temporary
                t[i:i + 64] = torch.randn(min(64, n - i), D, h, w, generator=g, device=device).half()
            out[k].append(t)
    if pin:
        out = {k: [t.pin_memory() for t in v] for k, v in out.items()}
    return out


def nbytes(d):
    return sum(t.numel() * t.element_size() for v in d.values() for t in v)


def views_of(inp, ops):
    an = ops.AnchorSpec(HW)
    head = ops.HeadViews.from_levels(inp["reg"], inp["obj"], inp["cls"], an)
    feats = tuple(ops.view_levels(inp[k]) for k in ("f_cls", "f_reg", "f_edge"))
    return head, feats


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU arm
def cpu_stage_sample(n_clips, seed=2024):
    """The reference's CPU path for the stage (oracle port of its fp32 PyTorch code), all host threads.
    Returns (clip_frames_per_s, seconds, threads)."""
    import oracle
    torch.set_num_threads(os.cpu_count() or 1)
    sd = oracle.init_stage_weights(C, dim=D, seed=seed)
    head, feats = oracle.synth_head_outputs(F, HW, C, dim=D, seed=seed)
    te = oracle.timing_signal_1d(torch.arange(LF), 256)
    kw = dict(selection="A", select_kwargs=dict(nms_thre=0.75, pre_k=PRE_K, top_k=TOP_K), nms_thresh=0.5)

    def one():
        dec = oracle.decode_outputs(head, HW, [8, 16, 32])          # decode_outputs belongs to the stage (tscd_head.py:378)
        oracle.stage_tscd(sd, dec, feats[0], feats[1], feats[2], te, C, LF, F - LF, **kw)

    one()                                                            # warm-up (also builds oracle/_build)
    t0 = time.perf_counter()
    for _ in range(n_clips):
        one()
    dt = time.perf_counter() - t0
    return n_clips * F / dt, dt, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    per_step = 2
    for _ in range(min(args.warmup, 1)):
        cpu_stage_sample(1)
    fps, dt, thr = cpu_stage_sample(per_step * args.steps)
    sample = f"{per_step * args.steps} clips x {F} frames, fp32, {thr} threads (oracle port of the reference's PyTorch CPU path)"
    line = {"impl": "reference", "metric": "clip-frames/sec of TSCD aggregation stage", "value": fps, "unit": "clip-frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_step": per_step},
            "cpu_baseline": {"value": fps, "unit": "clip-frames/s", "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "clip-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- roofline
def algorithmic_work(name, counts, B, need_reg_calls):
    """Algorithmic FLOPs / bytes per launch of the dominant kernel (formulas in DESIGN.md section 5)."""
    n_loc = sum(sum(counts[b * F:b * F + LF]) for b in range(B))
    pairs = 0
    for b in range(B):
        ng = sum(counts[b * F + LF:(b + 1) * F])
        pairs += sum(n * (n + ng) for n in counts[b * F:b * F + LF])
    if name == "tscd_attn_pv":        # QK^T of both branches (1024 flop/pair) + attn@v_cls (+ attn@v_reg)
        return "tensor", pairs * (1536 + 2048) / 2.0   # mean of the two launches per step (agg: 1536/pair, agg_iou: 2048/pair)
    if name == "tscd_attn_round2":    # raw-v cosine (512, +512 with the obj mask) + weights@V (512); mean over the 3 launches
        return "tensor", pairs * (1024 + 1024 + 1536) / 3.0
    if name == "tscd_select":         # obj plane + survivors' rows: A*s + P*(5+C)*s per frame
        return "hbm", B * F * (6804 * 2 + PRE_K * (5 + C) * 2 + PRE_K * 28)
    if name == "tscd_gather":
        return "hbm", B * F * (2 * TOP_K * 3 * D * 2 + TOP_K * (7 + C) * 4)
    if name == "tscd_linear":         # sum over the step's GEMMs of 2*M_valid*N*K, mean per launch
        return "tensor", sum(2.0 * m * N * K for (m, N, K) in LINEAR_SHAPES) / max(1, len(LINEAR_SHAPES))
    if name == "tscd_cafm_chain":     # SURVEY 8(d) K5, the part inside the recurrence: q projection + cosine attention
        fl = 0.0
        for b in range(B):
            for n in counts[b * F:b * F + LF]:
                fl += 2.0 * n * D * D + 4.0 * n * n * D
        return "tensor", fl
    return "tensor", None


LINEAR_SHAPES = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=64, help="clips per GPU per step (BASELINE.json configs[4]: 64 clips x 32 frames)")
    ap.add_argument("--e2e-clips", type=int, default=32)
    ap.add_argument("--e2e-chunk", type=int, default=8, help="clips per pipelined chunk of the host-buffer path")
    ap.add_argument("--cpu-clips", type=int, default=40, help="clips timed for cpu_baseline (rank 0, N=1)")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--split", type=int, default=1, help="concurrent sub-batches (streams) per step")
    ap.add_argument("--nchw-logits", action="store_true", help="head logits as NCHW planes instead of channels_last")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    global LOGITS_CHANNELS_LAST
    LOGITS_CHANNELS_LAST = not args.nchw_logits
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: tscd_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from tscd_b200 import _lib as L, ops, selection, stage, weights

    B = args.clips
    cfg = stage.StageConfig(num_classes=C, selection=selection.SelectionConfig(mode="A", pre_k=PRE_K, top_k=TOP_K, nms_thresh=0.75))
    st = stage.AggregationStage(cfg, weights.random_state_dict(C, D, seed=2024), device=dev)
    inp = synth_s1(B, dev, seed=2024 + rank)
    head, feats = views_of(inp, ops)
    te = torch.cat([weights.timing_signal_1d(torch.arange(LF), 256)] * B, 0).to(dev)

    # the step's clips are processed as `--split` independent sub-batches on concurrent streams (clips are independent
    # units; the kernels of one sub-batch fill the SMs another one leaves idle)
    nsp = max(1, min(args.split, B))
    bounds = [(B * i) // nsp for i in range(nsp + 1)]
    parts = []
    for i in range(nsp):
        c0, c1 = bounds[i], bounds[i + 1]
        sub = {k: [t[c0 * F:c1 * F] for t in v] for k, v in inp.items()}
        hd, ft = views_of(sub, ops)
        parts.append((hd, ft, te[c0 * LF:c1 * LF], c1 - c0))

    def step():
        if nsp == 1:
            return [st.forward(head, feats, torch.float16, te, B, F, LF)]
        return st.forward_concurrent(parts, torch.float16, F, LF)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up; then ONE sub-batch is launched alone on one stream and profiled per entry point (CUDA events around
    #      every C-ABI call) to find the dominant kernel: same launch shapes as in the timed region, no overlap ----
    for i in range(args.warmup):
        outs = step()
    torch.cuda.synchronize()
    for o, (_, _, _, nb) in zip(outs, parts):
        st.to_lists(o, nb, LF)                   # raises if the stage reported a capacity error
    hd0, ft0, te0, B0 = parts[0]
    L.profile = {"names": None, "events": {}}
    out = st.forward(hd0, ft0, torch.float16, te0, B0, F, LF)
    torch.cuda.synchronize()
    per_kernel = {k: sum(s.elapsed_time(e) for s, e in v) for k, v in L.profile["events"].items()}
    for (M, N, K, md) in L.profile.get("linear", []):
        LINEAR_SHAPES.append((M if md is None else min(M, int(md.item())), N, K))
    calls = {k: len(v) for k, v in L.profile["events"].items()}
    top = max(per_kernel, key=per_kernel.get)
    counts = out["sel"]["sel_count"].cpu().tolist()

    # launches per step (claim for `gpu_launches`) and host time of the eager launch sequence
    L.profile = None
    L.launch_count = 0
    h0 = time.perf_counter()
    outs = step()
    host_ms = 1e3 * (time.perf_counter() - h0)
    launches_per_step = L.launch_count
    torch.cuda.synchronize()

    # ---- CUDA graph of one step (the stage has no host sync); eager launches remain available with --no-graph ----
    graph = None
    if not args.no_graph:
        try:
            graph, outs = st.capture_fn(step)
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
        except Exception as e:   # noqa: BLE001
            sys.stderr.write(f"CUDA graph capture failed ({e}); timing eager launches\n")
            graph = None
            torch.cuda.synchronize()

    # ---- timed region ----
    L.profile = {"names": {top}, "events": {}} if (graph is None and nsp == 1) else None
    clocks = ClockSampler(local)
    clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            outs = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = launches_per_step * args.steps
    if graph is None and nsp == 1:
        top_ms = [s.elapsed_time(e) for s, e in L.profile["events"][top]]
    else:
        # inside a graph replay individual launches cannot be bracketed by events: the dominant kernel's duration is
        # measured live in the same process right after the timed region, launched eagerly on the same stream
        L.profile = {"names": {top}, "events": {}}
        for _ in range(3):
            st.forward(hd0, ft0, torch.float16, te0, B0, F, LF)     # one sub-batch alone: the launch shape of the timed region
        torch.cuda.synchronize()
        top_ms = [s.elapsed_time(e) for s, e in L.profile["events"][top]]
    L.profile = None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * F * args.steps / (ms / 1e3)

    # ---- e2e: host (pinned) inputs -> detections on the host, through AggregationStage.forward_host ----
    # Every step: H2D of the head logits (copy stream, chunk-pipelined with compute), zero-copy gather of the kept
    # proposals' feature rows out of the pinned feature planes, one D2H of the padded detections per chunk.
    Be = args.e2e_clips
    inp_mib = nbytes(inp) / 2**20
    launch_mode = "cuda_graph" if graph is not None else "eager"
    del graph, out, outs, parts, hd0, ft0, inp, head, feats
    torch.cuda.empty_cache()
    dev_src = synth_s1(Be, dev, seed=99 + rank)
    host = {k: [torch.empty(t.shape, dtype=t.dtype, pin_memory=True,
                                memory_format=torch.channels_last if (k.startswith("f_") or LOGITS_CHANNELS_LAST) else torch.contiguous_format).copy_(t)
                for t in v] for k, v in dev_src.items()}
    for k, v in host.items():
        for t, d in zip(v, dev_src[k]):
            assert t.is_pinned()
    del dev_src
    torch.cuda.empty_cache()
    te_e = torch.cat([weights.timing_signal_1d(torch.arange(LF), 256)] * Be, 0).pin_memory()

    def e2e_step():
        return st.forward_host(host, HW, te_e, Be, F, LF, chunk_clips=args.e2e_chunk)

    for _ in range(2):
        res, res_ori, h2d, d2h = e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        res, res_ori, h2d, d2h = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * Be * F * e2e_steps / e2e_s
    host_resident = nbytes(host)

    if rank != 0:
        return
    pk = peaks()
    bound, work = algorithmic_work(top, counts, B0, 1)
    avg_ms = statistics.mean(top_ms)
    roof = {"kernel": top, "bound": bound, "achieved": None, "peak": pk["hbm"] if bound == "hbm" else pk["tf_sust"],
            "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": None, "traffic": None,
            "avg_launch_ms": avg_ms, "launches_per_step": calls.get(top, 0), "peak_source": pk["src"] + ", sustained",
            "share_of_step": avg_ms * calls.get(top, 0) * nsp / (ms / args.steps),
            "note": f"per-launch figures are for one sub-batch of {B0} clips launched alone ({nsp} sub-batches per step run concurrently)",
            "per_entry_ms_one_step": {k: round(v, 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
            "calls_one_step": calls}
    if work is not None:
        ach = work / (avg_ms / 1e3) / (1e9 if bound == "hbm" else 1e12)
        roof["achieved"], roof["frac"] = ach, ach / roof["peak"]
    # DRAM traffic per launch of the dominant kernel from the committed `ncu` capture of the same workload
    # (profiles/traffic_r1.json, made by tools/profile_step.py --clips 64 under ncu); null for other batch sizes
    tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if os.path.exists(tpath) and B == 64 and nsp == 1:
        names = {"tscd_linear": "gemm_tn_kernel", "tscd_attn_round2": "attn_round2_kernel", "tscd_attn_pv": "attn_pv_kernel",
                 "tscd_attn_prep": "attn_prep_kernel", "tscd_select": "select_kernel|classmax_kernel", "tscd_nms": "nms_",
                 "tscd_gather": "rows_gather_kernel", "tscd_cafm_chain": "cafm_chain", "tscd_cafm_cost": "cafm_cost", "tscd_cafm_lap": "cafm_lap"}
        pats = names.get(top, top).split("|")
        tj = json.load(open(tpath))
        sel = [v for k, v in tj.items() if any(p_ in k for p_ in pats)]
        if sel:
            roof["traffic"] = sum(v["dram_bytes_per_launch"] * v["launches"] for v in sel) / max(1, calls.get(top, 1))
            roof["traffic_source"] = "profiles/traffic_r1.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, per C-ABI call)"
    line = {"metric": "clip-frames/sec of TSCD aggregation stage", "value": value, "unit": "clip-frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu_per_step": B, "concurrent_sub_batches": nsp, "seam": "S1 raw per-level conv outputs, fp16, " + ("channels_last logits and features (what the drop-in head's channels_last conv towers emit)" if LOGITS_CHANNELS_LAST else "NCHW logits, channels_last features"),
                       "l2": f"inputs per step ({inp_mib:.0f} MiB/GPU) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"clip-parallel x{world}, no collective"},
            "clocks": clk, "gpu_launches": launches, "launch_mode": launch_mode,
            "host_ms_per_eager_step": host_ms,
            "e2e": {"value": e2e_val, "unit": "clip-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "clips_per_gpu_per_step": Be, "chunk_clips": args.e2e_chunk, "steps": e2e_steps,
                    "host_resident_input_bytes_per_step": host_resident,
                    "note": "inputs are pinned HOST tensors; head logits are copied H2D, the 256-ch feature planes are read in place "
                            "(zero-copy gather of the kept rows); h2d counts both"},
            "roofline": roof}
    if world == 1:
        fps, dt, thr = cpu_stage_sample(args.cpu_clips)
        line["cpu_baseline"] = {"value": fps, "unit": "clip-frames/s", "cores": thr, "kind": "port",
                                "sample": f"{args.cpu_clips} clips x {F} frames of the same workload, fp32, {dt:.1f} s"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
