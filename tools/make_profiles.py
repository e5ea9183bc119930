"""Regenerate the round-2 files under profiles/ from the scratch captures in gpurun_out/ (tools/gpu_r2_prof.sh):
python tools/make_profiles.py"""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summary(rep):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout


# ---- ncu --set full summaries
parts = ["# r02 `ncu --set full --clock-control none --import-source on` summaries (B200, end of round 2)\n",
         "One capture per kernel, each taken after the same command had exited 0 without ncu (`tools/gpu_r2_prof.sh`; the plain",
         "runs' output is quoted under every table).  Durations under ncu are cold-cache and serialised; DRAM bytes and",
         "instruction counts are what these tables are for.  Raw reports: `gpurun_out/r02_*.ncu-rep` (scratch).\n"]
notes = {
    "scan_c2_chain": "C2 shape (100 k rows x 256 B), dependent chain, one scan per launch: DRAM read = 29.06 MB per scan = the algorithmic 28.9 MB (100 k x (256 + 33) B) -> no re-reads.",
    "scan_c4_chain": "C4 shape (1 M rows x 1 KB), one scan per launch: 1.056 GB read per scan = algorithmic (1 M x 1057 B).",
    "k1_c2": "K1 on C2 (150 MB of letters -> 25.6 MB of histograms): launch 0 = `validate_kernel` (load time), launch 1 = `kmer_count_kernel<1>` (letters -> 2-bit codes in registers -> counts, one pass).",
    "keys_c4": "K2b center-tile kernel on C4: 150 centers x 1 M rows; every row is read ONCE (1.09 GB of DRAM reads for 1.02 GB of histograms; round 1 read 154 GB).",
    "nw_c2": "K4 on 3000 pairs of 1.5 kb (C2 training batch), 16 rows per lane, DPX three-input maxima.",
    "phase_a_c2": "The persistent Phase-A kernel on C2 (`bin/meshclust`): ONE launch runs all 2001 scans, the bvec, means and nearest members; 31 MB of DRAM reads in total -- after the first scan the 25.6 MB matrix is served from L2 (hit rate 99 %), the step time is exchange latency, not bandwidth.",
}
for name in ["scan_c2_chain", "scan_c4_chain", "phase_a_c2", "k1_c2", "keys_c4", "nw_c2"]:
    rep = os.path.join(G, f"r02_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    parts.append(summary(rep))
    parts.append(notes[name] + "\n")
    plain = os.path.join(G, "r02_" + {"scan_c2_chain": "scan_c2", "scan_c4_chain": "scan_c4", "phase_a_c2": "pa_c2"}.get(name, name) + ".plain.log")
    if os.path.exists(plain):
        lines = [l.rstrip() for l in open(plain) if re.search(r"GCUPS|keys in|histograms|Accumulation|^\[\(", l)]
        if lines:
            parts.append("plain run (no profiler):\n```\n" + "\n".join(lines[-4:]) + "\n```\n")
open(os.path.join(P, "r02_ncu_full.md"), "w").write("\n".join(parts))

# ---- DRAM traffic of one scan for bench.py's roofline.traffic
traffic = {}
for key, name in (("c2_256", "scan_c2_chain"), ("c4_1024", "scan_c4_chain")):
    rep = os.path.join(G, f"r02_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, data = raw(rep)
    def val(row, k):
        i = hdr.index(k)
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(row[i]) * mult
    per = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in data]
    traffic[key] = {"dram_bytes_per_launch": int(sum(per) / len(per)),
                    "kernel": data[0][hdr.index("Kernel Name")].split("(")[0],
                    "source": f"profiles/r02_ncu_full.md: gpurun_out/r02_{name}.ncu-rep (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(per)} launches of the dependent chain)"}
if traffic:
    json.dump(traffic, open(os.path.join(P, "scan_traffic.json"), "w"), indent=1)

# ---- launch list of the bench command
src = os.path.join(G, "r02_bench_launches.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if l.startswith('"')]
    open(os.path.join(P, "r02_bench_launches.csv"), "w").writelines(lines)
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    tot = sum(a[1] for a in agg.values())
    out = ["# r02 launch list of `python bench.py --steps 1 --warmup 3 --scaling-only` (ncu --metrics gpu__time_duration.sum --clock-control none -c 700)\n",
           "per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes\n",
           "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for k, (c, t) in agg.items():
        out.append(f"| `{k}` | {c} | {t:.1f} | {t / c:.2f} | {100 * t / tot:.1f}% |")
    out.append("\nThe timed region launches only `scan_tma_kernel<1, 256, 0, ...>` (one launch per dependent scan; the batched form carries the")
    out.append("independent-scans micro-benchmark): its share of a timed step is 100 %.  Everything else in the list is set-up (K1, model fit,")
    out.append("replica load).  `-c 700` cuts the list inside the warm-up; the plain run of the same command printed:\n")
    pl = os.path.join(G, "r02_bench.plain.log")
    if os.path.exists(pl):
        js = [l for l in open(pl) if l.startswith("{")]
        if js:
            d = json.loads(js[-1])
            out.append("```\n" + json.dumps({k: d[k] for k in ("metric", "value", "ms_per_step", "gpu_launches", "roofline") if k in d}, indent=1) + "\n```")
    open(os.path.join(P, "r02_bench_launches.md"), "w").write("\n".join(out) + "\n")

# ---- SASS
sass = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_excerpt.py")], capture_output=True, text=True).stdout
open(os.path.join(P, "r02_sass_excerpt.md"), "w").write(sass)
print("profiles/ regenerated")
