"""Summarise an ncu report (raw page) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    g = lambda k: d.get(k, ("n/a", ""))[0]
    out.append(f"## {g('Kernel Name')}  grid {g('Grid Size')} block {g('Block Size')}")
    keys = [
        ("duration", "gpu__time_duration.sum"), ("registers/thread", "launch__registers_per_thread"),
        ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
        ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("fma pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("fmaheavy pipe %", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active"),
        ("alu pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("lsu pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("xu pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("sm throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
        ("dram throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 hit %", "lts__t_sector_hit_rate.pct"), ("L1 hit %", "l1tex__t_sector_hit_rate.pct"),
        ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("warp instructions", "smsp__inst_executed.sum"),
    ]
    for name, k in keys:
        v, u = d.get(k, ("n/a", ""))
        out.append(f"- {name}: {v} {u}")
    out.append("- stalls per issue: " + ", ".join(
        f"{h.split('issue_stalled_')[1].split('_per_issue')[0]} {float(v):.2f}" for h, v in
        sorted(((h, d[h][0]) for h in d if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")),
               key=lambda t: -float(t[1]))[:8]))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(f"# ncu summary of {rep}\n\n" + text + "\n")
