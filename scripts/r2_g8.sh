#!/usr/bin/env bash
# round 2, 8-GPU call: strong + weak scaling lines at N = 2, 4, 8, the host-link ceiling with all ranks active, the 2-rank learner test
set -u
mkdir -p gpurun_out
python tools/actorbench.py --paths default --launches 4 > gpurun_out/g_actor.log 2>&1
python -m pytest tests/test_gpu_ddpg.py -m gpu -q -k two_rank > gpurun_out/g_two_rank.log 2>&1
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g_bench_${N}gpu.json 2> gpurun_out/g_bench_${N}gpu.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/pciebench_multi.py > gpurun_out/g_pcie8.log 2>&1
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/g_bench_1gpu.json 2> gpurun_out/g_bench_1gpu.err
nvidia-smi topo -m > gpurun_out/g_topo.txt 2>&1
cat gpurun_out/g_actor.log; tail -2 gpurun_out/g_two_rank.log
for N in 1 2 4 8; do python - <<EOF
import json
try:
    d = json.loads(open("gpurun_out/g_bench_${N}gpu.json").readline())
    w = d.get("weak_scaling_1M_envs_per_gpu", {})
    print($N, "strong", "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"] * 1e3), "e2e %.3e" % d["e2e"]["value"],
          "weak", "%.3e" % w.get("value", 0), "nograph", d.get("same_steps_without_cuda_graph", {}).get("ms_per_step"))
except Exception as e:
    print($N, "error", e)
EOF
done
tail -3 gpurun_out/g_pcie8.log | cut -c1-1500
