"""Measure the host link of the GPU box: pinned H2D alone, D2H alone, and both at once on two streams.
The e2e number of bench.py (VecMREnv.step_host) is bounded by these.  GPU box only."""
import json

import torch

MB = 64
n = MB * (1 << 20) // 8
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=20):
    torch.cuda.synchronize()
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


for _ in range(2):
    run(True, True, 3)
t_h2d, t_d2h, t_both = run(True, False), run(False, True), run(True, True)
gb = MB * (1 << 20) / 1e9
out = {"buffer_MB": MB, "h2d_GBps": gb / (t_h2d * 1e-3), "d2h_GBps": gb / (t_d2h * 1e-3),
       "both_ms": t_both, "both_total_GBps": 2 * gb / (t_both * 1e-3), "serial_sum_ms": t_h2d + t_d2h}
print(json.dumps(out))
json.dump(out, open("gpurun_out/pcie.json", "w"), indent=1)
