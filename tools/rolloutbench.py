import os, sys
sys.path.insert(0, os.getcwd())
sys.argv=["stepbench"]
import tools.stepbench as sb, torch
for sigma in (0.0, 1.0):
    ms, eps = sb.time_rollout(1 << 20, torch.float64, sigma, 64)
    print(f"rollout K=64 sigma={sigma} {ms:8.3f} ms {eps/1e9:8.2f} G/s")
