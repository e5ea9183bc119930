"""In-tree build of the CUDA library (sm_100a only) and the host CLI.

    python -m meshclust_b200.build            # libmeshclust_b200.so (+ bin/meshclust when host sources exist)

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
# MC_LIB_VARIANT=name + MC_EXTRA_NVCC_FLAGS="-D..." build an experimental variant next to the product
# library (libmeshclust_b200_<name>.so, selected at run time with MESHCLUST_B200_LIB)
_VARIANT = os.environ.get("MC_LIB_VARIANT", "")
LIB = os.path.join(PKG, f"libmeshclust_b200{'_' + _VARIANT if _VARIANT else ''}.so")
BIN = os.path.join(ROOT, "bin", "meshclust")
OBJ = os.path.join(PKG, "_build" + ("_" + _VARIANT if _VARIANT else ""))

CU_SOURCES = ["capi.cu", "fasta_ingest.cu", "kmer_hist.cu", "pair_kernels.cu", "scan_tma.cu", "center_mean.cu", "phase_a.cu", "nw_identity.cu", "peer_exchange.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                 # FP64 parity: the only fused ops are the explicit fma() calls
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
] + os.environ.get("MC_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found; meshclust_b200 has no CPU fallback")
    return nv


def _stamp(paths) -> str:
    h = hashlib.sha1()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _deps(dirname):
    return [os.path.join(dirname, f) for f in os.listdir(dirname)
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp"))]


def build_lib(force: bool = False, verbose: bool = False) -> str:
    deps = _deps(CSRC) + [os.path.join(ROOT, "include", "meshclust_b200.h")]
    stamp = _stamp(deps)
    stamp_file = os.path.join(OBJ, "lib.stamp")
    if (not force and os.path.exists(LIB) and os.path.exists(stamp_file)
            and open(stamp_file).read() == stamp):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nv = _nvcc()
    objs = []
    procs = []
    for src in CU_SOURCES:
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nv, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[build] {src} FAILED\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"[build] {src}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nv, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs])
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


def build_cli(force: bool = False) -> str | None:
    """bin/meshclust: host C++ (reference CLI re-created) linked against the CUDA library."""
    main_cpp = os.path.join(HOST, "main.cpp")
    if not os.path.exists(main_cpp):
        return None
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp")]
    deps = _deps(HOST) + [os.path.join(ROOT, "include", "meshclust_b200.h")]
    stamp = _stamp(deps)
    stamp_file = os.path.join(OBJ, "cli.stamp")
    if (not force and os.path.exists(BIN) and os.path.exists(stamp_file)
            and open(stamp_file).read() == stamp and os.path.getmtime(BIN) >= os.path.getmtime(LIB)):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-Wall", "-Wno-sign-compare",
           "-I", os.path.join(ROOT, "include"), *srcs, "-o", BIN,
           "-L", PKG, "-lmeshclust_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../meshclust_b200"]
    subprocess.check_call(cmd)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return BIN


def build_all(force: bool = False, verbose: bool = False):
    lib = build_lib(force, verbose)
    cli = None if _VARIANT else build_cli(force)
    return lib, cli


if __name__ == "__main__":
    lib, cli = build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
    if cli:
        print(cli)
