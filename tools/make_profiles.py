#!/usr/bin/env python
"""Turn the ncu CSV logs of one profiled step into the summaries committed under profiles/.

On the GPU box (one step of the bench workload, bracketed by cudaProfilerStart/Stop in tools/profile_step.py):

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
      --log-file gpurun_out/launches_r2.csv python tools/profile_step.py --clips 64
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,\
sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,\
launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__registers_per_thread,\
launch__shared_mem_per_block_dynamic --clock-control none --profile-from-start off -c 80 --csv \
      --log-file gpurun_out/ncu_metrics_r2.csv python tools/profile_step.py --clips 64

Here:  python tools/make_profiles.py   ->  profiles/launches_r2.csv, profiles/ncu_kernels_r2.csv, profiles/traffic_r2.json
"""
import collections
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")


def main():
    shutil.copy(os.path.join(SRC, "launches_r2.csv"), os.path.join(DST, "launches_r2.csv"))
    rows = [r for r in csv.reader(open(os.path.join(SRC, "ncu_metrics_r2.csv"))) if len(r) > 5]
    hdr = rows[0]
    ki, ii, mi, vi, ui, gi = (hdr.index(k) for k in ("Kernel Name", "ID", "Metric Name", "Metric Value", "Metric Unit", "Grid Size"))
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[ii], {"name": r[ki], "grid": r[gi]})
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[mi].endswith("bytes.sum") or r[mi].endswith("_read.sum") or r[mi].endswith("_write.sum"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1)
        d[r[mi]] = v
    out = [["id", "kernel", "grid", "time_us", "dram_read_MB", "dram_write_MB", "dram_GBps", "l2_MB", "sm_throughput_pct",
            "tensor_pipe_active_pct", "xu_pipe_pct", "warps_active_pct", "regs", "dyn_smem_KB", "occ_limit_smem", "occ_limit_regs"]]
    agg = collections.OrderedDict()
    for i, d in launches.items():
        t, rd, wr = d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
        name = d["name"].split("(")[0].replace("void ", "").replace("tscd::", "")
        out.append([i, name, d["grid"], f"{t:.1f}", f"{rd / 1e6:.1f}", f"{wr / 1e6:.1f}", f"{(rd + wr) / max(t, 1e-9) / 1e3:.0f}",
                    f"{d.get('lts__t_bytes.sum', 0) / 1e6:.1f}", f"{d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f}",
                    f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f}",
                    f"{d.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 0):.1f}",
                    f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.1f}", int(d.get("launch__registers_per_thread", 0)),
                    f"{d.get('launch__shared_mem_per_block_dynamic', 0) / 1e3:.1f}", int(d.get("launch__occupancy_limit_shared_mem", 0)),
                    int(d.get("launch__occupancy_limit_registers", 0))])
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += rd + wr
    csv.writer(open(os.path.join(DST, "ncu_kernels_r2.csv"), "w")).writerows(out)
    tot = sum(a[1] for a in agg.values())
    traffic = {}
    print(f"{'kernel':40s} {'n':>3s} {'us':>8s} {'share':>6s} {'MB/launch':>9s} {'GB/s':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] > 5:
            print(f"{k[:40]:40s} {a[0]:3d} {a[1]:8.1f} {100 * a[1] / tot:5.1f}% {a[2] / a[0] / 1e6:9.1f} {a[2] / a[1] / 1e3:7.0f}")
        traffic[k] = {"launches": a[0], "dram_bytes_per_launch": a[2] / a[0], "time_us_total": a[1]}
    print(f"sum of kernel times {tot:.1f} us over {len(launches)} launches")
    # DRAM bytes per launch keyed like bench.py's kernel rows (launch order of the serialised pass: agg_iou first, then agg)
    per_kernel = collections.OrderedDict()
    for i, d in launches.items():
        name = d["name"].split("(")[0].replace("void ", "").replace("tscd::", "")
        per_kernel.setdefault(name.split("<")[0], []).append(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
    rules = {"tscd_select": ("select_rows_kernel", 0), "tscd_gather": ("rows_gather_kernel", 0), "tscd_attn_pv:agg_iou": ("attn_pv_kernel", 0),
             "tscd_attn_pv:agg": ("attn_pv_kernel", 1), "tscd_attn_round2:agg_iou.cls": ("attn_round2_kernel", 0),
             "tscd_attn_round2:agg_iou.obj": ("attn_round2_kernel", 1), "tscd_attn_round2:agg.cls": ("attn_round2_kernel", 2),
             "tscd_cafm_chain": ("cafm_chain_fast_kernel", 0), "tscd_cafm_prep": ("cafm_prep_kernel", 0), "tscd_cafm_cost": ("cafm_cost16_kernel", 0),
             "tscd_frame_attention": ("frame_attention16_kernel", 0), "tscd_residual_ln2": ("residual_ln2_kernel", 0),
             "tscd_nms:pre": ("nms_kernel", 0), "tscd_nms:final_det": ("nms_matrix_kernel", 0)}
    by_key = {k: per_kernel[n][j] for k, (n, j) in rules.items() if n in per_kernel and j < len(per_kernel[n])}
    clips = int(os.environ.get("TSCD_PROFILE_CLIPS", "148"))
    json.dump({"clips_per_replay": clips, "by_bench_key": by_key, "by_cuda_kernel": traffic}, open(os.path.join(DST, "traffic_r2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
