#!/usr/bin/env bash
# round 2, short 8-GPU refresh: strong + weak scaling lines at N = 8, 4, 2 and the 2-rank learner test (the host-link
# measurement of scripts/r2_g8.sh is not repeated)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ddpg.py -m gpu -q -k two_rank > gpurun_out/g_two_rank.log 2>&1
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g_bench_${N}gpu.json 2> gpurun_out/g_bench_${N}gpu.err
done
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/g_bench_1gpu.json 2> gpurun_out/g_bench_1gpu.err
tail -2 gpurun_out/g_two_rank.log
for N in 1 2 4 8; do python - <<EOF
import json
try:
    d = json.loads(open("gpurun_out/g_bench_${N}gpu.json").readline())
    w = d.get("weak_scaling_1M_envs_per_gpu", {})
    f = d.get("fused_rollout_k64", {})
    print($N, "strong", "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"] * 1e3), "e2e %.3e" % d["e2e"]["value"],
          "weak", "%.3e" % w.get("value", 0), "fused", "%.3e" % f.get("value", 0), "nograph", d.get("same_steps_without_cuda_graph", {}).get("ms_per_step"))
except Exception as e:
    print($N, "error", e)
EOF
done
