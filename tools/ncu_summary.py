"""Summarise an .ncu-rep (raw page) into a small markdown table: python tools/ncu_summary.py rep [rep...]"""
import csv, io, subprocess, sys
WANT = [
 ("gpu__time_duration.sum", "duration"),
 ("dram__bytes_read.sum", "DRAM read"),
 ("dram__bytes_write.sum", "DRAM write"),
 ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
 ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
 ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
 ("launch__registers_per_thread", "registers/thread"),
 ("launch__grid_size", "grid"),
 ("launch__block_size", "block"),
 ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
 ("smsp__inst_executed.sum", "warp instructions"),
 ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
 ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"### {rep.split('/')[-1]}\n")
    kn = hdr.index("Kernel Name")
    print("kernel: `" + data[0][kn][:90] + "`\n")
    print("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    print("|---|" + "---|" * len(data))
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            print(f"| {label} ({units[i]}) | " + " | ".join(r[i] for r in data) + " |")
    stall = [(h, i) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h]
    tot = sum(float(data[-1][i] or 0) for _, i in stall)
    top = sorted(((float(data[-1][i] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h, i in stall), reverse=True)[:6]
    print("\nstall samples (last launch): " + ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in top if tot) + "\n")
