"""Unit self-checks of the host-side containers (CPU): compiled from tests/units/*.cpp."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bvec_selfcheck(tmp_path):
    # binary-search index_of == the reference's literal loop (bvec.cpp:123-149); bitmap-backed bins ==
    # an erase-based model: pop order, sizes, and "a bvec range is a contiguous alive row range"
    # (also: insert + finalize against the literal rule of bvec.cpp:152-177, 209-218; built with and without OpenMP:
    # the bin choice uses `omp simd` reductions, the sorts OpenMP tasks)
    for flags in ([], ["-fopenmp"]):
        exe = os.path.join(str(tmp_path), "bvec_selfcheck" + ("_omp" if flags else ""))
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", *flags, os.path.join(ROOT, "tests", "units", "bvec_selfcheck.cpp"), "-o", exe])
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def test_fasta_parallel_parser_selfcheck(tmp_path):
    # the multi-threaded FASTA parser == the serial one (ChromListMaker.cpp:92-120 semantics) on regular
    # files; CRLF files, headers without sequence and other irregular inputs fall back to the serial path
    exe = os.path.join(str(tmp_path), "fasta_selfcheck")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", os.path.join(ROOT, "tests", "units", "fasta_selfcheck.cpp"), "-o", exe])
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env={**os.environ, "OMP_NUM_THREADS": "7"})
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def test_lazy_sort_selfcheck(tmp_path):
    # Trainer::split's ~150 sorts of all points run lazily (host/lazy_sort.hpp): every position resolved must hold the
    # element libstdc++'s std::sort puts there -- ties included -- also through introsort's heapsort branch
    exe = os.path.join(str(tmp_path), "lazy_sort_selfcheck")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", os.path.join(ROOT, "tests", "units", "lazy_sort_selfcheck.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
