"""Host-link bandwidth with ALL ranks active at once (torchrun, one rank per GPU): pinned H2D, D2H and both directions,
per rank and summed — the ceiling of bench.py's multi-GPU e2e numbers.  Then the same after binding every rank to the
CPUs of its GPU's NUMA node and re-allocating the pinned buffers there (first touch after the bind).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pciebench_multi.py
"""
import json
import os
import subprocess

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = 64
n = MB * (1 << 20) // 8


def numa_of_gpu(idx):
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(idx)],
                             capture_output=True, text=True).stdout.strip().lower()
        bdf = bdf[-12:] if len(bdf) > 12 else bdf                # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        cpus = open(f"/sys/devices/system/node/node{max(node, 0)}/cpulist").read().strip()
        return node, cpus
    except Exception as e:
        return None, str(e)[:80]


def parse_cpulist(s):
    out = set()
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out.update(range(int(a), int(b) + 1))
        elif part.strip().isdigit():
            out.add(int(part))
    return out


def measure(tag):
    h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
    h_out = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(0.0)
    d_in = torch.empty(n, dtype=torch.float64, device="cuda")
    d_out = torch.ones(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=20):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    run(True, True, 3)
    gb = MB * (1 << 20) / 1e9
    res = {}
    for name, (a, b) in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True))):
        ms = run(a, b)
        mine = (2 if name == "both" else 1) * gb / (ms * 1e-3)
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = mine
        if world > 1:
            dist.all_reduce(t)
        res[name] = {"per_rank_GBps": [round(float(x), 1) for x in t.tolist()], "sum_GBps": round(float(t.sum()), 1)}
    return res


node, cpus = numa_of_gpu(local)
info = {"rank": rank, "gpu_numa_node": node, "node_cpus": cpus, "affinity_before": len(os.sched_getaffinity(0))}
before = measure("default")
bound = False
try:
    want = parse_cpulist(cpus) & os.sched_getaffinity(0) if node is not None else set()
    if want:
        os.sched_setaffinity(0, want)
        bound = True
except Exception as e:
    info["bind_error"] = str(e)[:80]
after = measure("bound")
gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, info)
else:
    gathered = [info]
if rank == 0:
    out = {"world": world, "buffer_MB": MB, "ranks": gathered, "default_placement": before,
           "numa_bound_placement": after, "bound": bound}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/pcie_multi_{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
