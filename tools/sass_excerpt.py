"""SASS evidence for profiles/: per kernel of the shipped library, how often the instructions that carry the
design appear (bulk copies + mbarrier waits, packed-byte SIMD, DPX, warp reductions, shared-memory reductions).
python tools/sass_excerpt.py > profiles/r02_sass_excerpt.md"""
import collections, re, subprocess, sys, os
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "meshclust_b200", "libmeshclust_b200.so")
KEEP = ["scan_tma_kernel", "phase_a_kernel", "kmer_count_kernel", "nw_kernel", "dist_keys_tile_kernel", "dist_keys_kernel", "accumulate_tail_kernel", "pair_list_kernel"]
PAT = re.compile(r"\b(UBLKCP[.\w]*|SYNCS[.\w]*|VABSDIFF4[.\w]*|IDP\.4A[.\w]*|VIMNMX3?[.\w]*|VIADDMNMX[.\w]*|REDUX[.\w]*|CREDUX[.\w]*|ATOMS[.\w]*|RED\.[.\w]*|ATOMG[.\w]*|ACQBULK|PREEXIT|MATCH[.\w]*|DFMA|DMUL|DADD|MUFU[.\w]*|SHF[.\w]*|LDGSTS[.\w]*|FENCE[.\w]*|MEMBAR[.\w]*)\b")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
cnt, total, arch = collections.defaultdict(collections.Counter), collections.Counter(), set()
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if fn and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        total[fn] += 1
        m = PAT.search(line)
        if m:
            cnt[fn][m.group(1)] += 1
print("# SASS excerpt of `meshclust_b200/libmeshclust_b200.so` (cuobjdump -sass; arch " + ", ".join(sorted(arch)) + ")\n")
print("Instruction counts per kernel instantiation (static, not executed counts).  UBLKCP = `cp.async.bulk` (1-D TMA),")
print("SYNCS.* = mbarrier arrive / try_wait, VABSDIFF4 + IDP.4A = packed-byte |p-q| and p.q reductions, VIMNMX3 / VIADDMNMX = DPX,")
print("REDUX / CREDUX = warp-wide reductions, ATOMS / RED = shared / global reductions, DFMA.. = the FP64 epilogue.\n")
for f in sorted(total):
    if not any(k in f for k in KEEP):
        continue
    d = demangle(f)
    if ("scan_tma" in d or "phase_a" in d or "dist_keys" in d or "pair_list" in d or "tail" in d) and not re.search(r"<1, (256|1024|4096)[,>]|<1>|<1, \(int\)", d):
        continue   # the u8 shapes of the BASELINE configs (k = 4, 5, 6)
    d = re.sub(r"\(.*", "", d)
    print(f"### `{d}`  ({total[f]} instructions)\n")
    print(", ".join(f"{k} x{v}" for k, v in sorted(cnt[f].items(), key=lambda kv: (-kv[1], kv[0]))) or "(none of the listed)")
    print()
