"""Turn ncu outputs brought back in gpurun_out/ into the markdown summaries kept under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/r01_v3_launches.csv  > profiles/r01_v3_launches.md
    python scripts/summarize_ncu.py full gpurun_out/r01_v3_prof_gemm.ncu-rep > profiles/r01_v3_gemm_ncu.md
    python scripts/summarize_ncu.py traffic gpurun_out/r02_ncu_gemm.ncu-rep gemm_bf16_tn_2cta > profiles/r02_gemm_traffic.json
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("time", "gpu__time_duration.sum"), ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("tensor pipe active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("L2 read sectors (from SMs)", "lts__t_sectors_srcunit_tex_op_read.sum"),
    ("dram throughput", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps active", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs/thread", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("cluster", "launch__cluster_size"), ("instructions", "smsp__inst_executed.sum"),
    ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("issue slots busy", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("stall: short scoreboard / issue", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall: long scoreboard / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall: barrier / issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall: math pipe throttle / issue", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall: not selected / issue", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("occupancy limit: registers", "launch__occupancy_limit_registers"), ("occupancy limit: shared memory", "launch__occupancy_limit_shared_mem"),
]


def short(name):
    return re.sub(r"^void |rs::|\(.*$", "", name)


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, seq = collections.OrderedDict(), []
    for r in rows[1:]:
        n, v = short(r[ki]), float(r[vi].replace(",", "")) / 1000.0
        seq.append((n, v))
        t = tot.setdefault(n, [0, 0.0])
        t[0] += 1
        t[1] += v
    total = sum(v for _, v in seq)
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {v:.0f} | {v / n:.1f} | {100 * v / total:.1f}% |")
    print(f"\nTotal {total / 1000:.1f} ms over {len(seq)} launches.")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = [f"{i}: {short(r[ki])}" for i, r in enumerate(data)]
    print("| metric | unit | " + " | ".join(cols) + " |\n|---|---|" + "---:|" * len(cols))
    for label, m in METRICS:
        if m not in hdr:
            continue
        j = hdr.index(m)
        print(f"| {label} (`{m}`) | {units[j]} | " + " | ".join(r[j] for r in data) + " |")


def traffic(path, pattern):
    """JSON for bench.py's roofline.traffic: mean dram read + write bytes per launch of the kernels matching `pattern`."""
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]] for r in data if re.search(pattern, r[ki])]
    print(json.dumps({"dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals), "kernel": pattern,
                      "source": f"ncu --set full capture {path.split('/')[-1]} (dram__bytes_read.sum + dram__bytes_write.sum, mean over {len(vals)} launches); "
                                "table in the profiles/ file of the same round"}))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        {"launches": launches, "full": full}[mode](sys.argv[2])
