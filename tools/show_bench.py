#!/usr/bin/env python
"""Pretty-print the per-kernel table of a bench.py JSON line:  python tools/show_bench.py gpurun_out/b.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ks = d.pop("kernels", [])
r = d["roofline"]
print(f"value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  replay {r['replay_ms']*1e3:.0f} us  serialised kernels {r['serialised_kernel_ms_per_replay']*1e3:.0f} us"
      f"  e2e {d.get('e2e', {}).get('value')}")
print(f"dominant: {r['kernel']} {r['bound']} frac {r['frac']}")
for k in ks:
    ach = "" if k["achieved"] is None else f"{k['achieved']:8.1f} {k['unit']:8s} frac {k['frac']:.3f}"
    print(f"{k['kernel']:38s} n={k['launches_per_step']} {k['avg_launch_ms']*1e3:8.1f} us {k['bound']:8s} {ach}  share {k['share_of_serialised_step']:.3f}")
