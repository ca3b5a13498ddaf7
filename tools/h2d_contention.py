#!/usr/bin/env python
"""Host->device copy bandwidth per GPU with 1..W ranks copying concurrently (names the e2e scaling limiter with data).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/h2d_contention.py

Every rank owns a pinned 1 GiB host buffer.  For k = 1, 2, 4, 8 the first k ranks copy it to their GPU 8 times back to back
(cudaMemcpyAsync on one stream, timed with CUDA events) while the others idle; rank 0 prints GB/s per active GPU and the sum.
Also prints the PCI bus id -> NUMA node of every GPU (/sys/bus/pci/devices/*/numa_node) and `nvidia-smi topo -m`."""
import json
import os
import subprocess

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 30
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.fill_(rank)
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    busid = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
    numa = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        busid = bus
        p = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
        numa = open(p).read().strip() if os.path.exists(p) else None
    except Exception as e:      # noqa: BLE001
        numa = f"n/a ({type(e).__name__})"
    results = {}
    k = 1
    while k <= world:
        dist.barrier()
        torch.cuda.synchronize()
        gbs = 0.0
        if rank < k:
            for _ in range(2):
                dst.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                dst.copy_(host, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            gbs = 8 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
        allg = [None] * world
        dist.all_gather_object(allg, gbs)
        results[k] = [round(x, 1) for x in allg[:k]]
        k *= 2
    info = [None] * world
    dist.all_gather_object(info, (busid, numa))
    if rank == 0:
        out = {"h2d_GBps_per_active_gpu": results, "sum_GBps": {k: round(sum(v), 1) for k, v in results.items()},
               "gpu_pci_numa": info, "host_cpus": os.cpu_count()}
        print(json.dumps(out))
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout)
            print(subprocess.run(["bash", "-c", "lscpu | grep -i 'numa\\|socket\\|model name'; free -g | head -2"], capture_output=True, text=True, timeout=20).stdout)
        except Exception:      # noqa: BLE001
            pass
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
