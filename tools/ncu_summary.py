"""Dev tool: summarise an .ncu-rep (raw metrics + executed-opcode histogram + top stall lines)."""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'lts__t_sector_hit_rate.pct']


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main(path, envs_per_warp=None):
    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        print("kernel:", r[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr:
                print(f"   {w:75s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
        st = [(float(r[i]), h) for i, h in enumerate(hdr)
              if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct') and r[i]]
        for v, h in sorted(st, reverse=True)[:7]:
            print('      stall', h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), round(v, 1))
    src = list(csv.reader(io.StringIO(run(["ncu", "-i", path, "--page", "source", "--csv"]))))
    hi = next(i for i, r in enumerate(src) if r and r[0] == 'Address')
    h2 = src[hi]
    iS, iE, iW = h2.index('Source'), h2.index('Instructions Executed'), h2.index('Warp Stall Sampling (All Samples)')
    ops, total = collections.Counter(), 0
    body = [r for r in src[hi + 1:] if len(r) > max(iW, iE, iS) and r[iE].isdigit()]
    for r in body:
        m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[iS])
        ops[m.group(2) if m else '?'] += int(r[iE]); total += int(r[iE])
    print("total warp instructions:", total)
    print("  ".join(f"{op}:{n}" for op, n in ops.most_common(30)))
    for r in sorted(body, key=lambda r: -int(r[iW]) if r[iW].isdigit() else 0)[:14]:
        print(f"   stall {r[iW]:>5s} exec {r[iE]:>8s}  {r[iS][:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
