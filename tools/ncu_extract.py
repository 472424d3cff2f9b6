#!/usr/bin/env python
"""Extract the judged metrics of every kernel in an .ncu-rep into a small CSV (profiles/) and print a digest."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ["Kernel Name", "gpu__time_duration", "launch__", "dram__bytes", "dram__throughput", "pipe_fma", "pipe_alu", "fmaheavy", "warps_active", "issue_active",
        "inst_executed.sum", "inst_executed.avg.per_cycle", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "local_op", "cycles_elapsed.avg", "issue_stalled", "clock_rate",
        "thread_inst_executed_per_inst"]
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel_index", "metric", "unit", "value"])
    for ki, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        for h, u in zip(hdr, units):
            if any(k in h for k in KEEP):
                w.writerow([ki, h, u, d[h]])
        g = lambda k: d.get(k, "")
        stalls = sorted(((float(v), h.split("stalled_")[1].split("_per")[0]) for h, v in d.items() if "issue_stalled" in h and "per_issue_active" in h and v not in ("", "n/a")), reverse=True)[:5]
        print(f"[{ki}] {g('Kernel Name')[:90]}\n    time {g('gpu__time_duration.sum')} ms  grid {g('launch__grid_size')} x {g('launch__block_size')}  regs {g('launch__registers_per_thread')}"
              f"  fmaheavy {g('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed')}%  alu {g('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed')}%"
              f"  warps_active {g('sm__warps_active.avg.pct_of_peak_sustained_active')}%  lanes/inst {g('smsp__thread_inst_executed_per_inst_executed.ratio')}"
              f"\n    dram rd {g('dram__bytes_read.sum')} wr {g('dram__bytes_write.sum')}  stalls {[(round(a, 2), b) for a, b in stalls]}")
