"""The host control flow of bin/meshclust (sampling, GLM, bvec, accumulate / update / merge
bookkeeping, CLSTR writer) against CLSTR files produced by the reference binary at --threads 1.

CPU (not gpu): the host sources are linked against tests/mock/mock_capi.cpp, a CPU stand-in for
the C-ABI built on the oracle -- test infrastructure only, the product binary never links it.
GPU: the real bin/meshclust (linked against libmeshclust_b200.so) on the same inputs."""
import os
import subprocess

import pytest

import _hostcases as H

ROOT = H.ROOT
HOST = os.path.join(ROOT, "meshclust_b200", "host")
MOCK_BIN = os.path.join(ROOT, "tests", "_build", "meshclust_hostlogic")


@pytest.fixture(scope="session")
def mock_cli():
    """Always reflects the current sources: the rebuild is keyed on a hash of their contents and of the
    flags (not on mtimes, which say nothing after a fresh clone)."""
    import hashlib
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp")]
    srcs += [os.path.join(ROOT, "tests", "mock", "mock_capi.cpp")]
    deps = srcs + [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".hpp")]
    deps += [os.path.join(ROOT, "oracle", "mc_oracle.c"), os.path.join(ROOT, "oracle", "mc_oracle.h"), os.path.join(ROOT, "include", "meshclust_b200.h")]
    cflags = ["-O2", "-march=x86-64-v3", "-ffp-contract=off", "-std=c11", "-fopenmp"]
    cxxflags = ["-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-Wno-sign-compare"]
    h = hashlib.sha1(" ".join(cflags + cxxflags).encode())
    for d in deps:
        h.update(d.encode())
        h.update(open(d, "rb").read())
    stamp_file = MOCK_BIN + ".stamp"
    if not (os.path.exists(MOCK_BIN) and os.path.exists(stamp_file) and open(stamp_file).read() == h.hexdigest()):
        os.makedirs(os.path.dirname(MOCK_BIN), exist_ok=True)
        obj = os.path.join(os.path.dirname(MOCK_BIN), "mc_oracle.o")
        subprocess.check_call(["/usr/bin/gcc", *cflags, "-c", os.path.join(ROOT, "oracle", "mc_oracle.c"), "-o", obj])
        subprocess.check_call(["/usr/bin/g++", *cxxflags, "-I", os.path.join(ROOT, "include"), *srcs, obj, "-o", MOCK_BIN, "-lm"])
        with open(stamp_file, "w") as f:
            f.write(h.hexdigest())
    return MOCK_BIN


def _run(binary, name, tmp_path, extra=(), env=None):
    paths, args = H.make_inputs(name, str(tmp_path))
    out = os.path.join(str(tmp_path), "out.clstr")
    r = subprocess.run([binary, *paths, *args, *extra, "--output", out], capture_output=True, text=True,
                       env=None if env is None else {**os.environ, **env})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return open(out, "rb").read(), r.stdout


@pytest.mark.parametrize("name", H.CPU_CASES)
def test_clstr_identical_to_reference_cpu(mock_cli, tmp_path, name):
    got, _ = _run(mock_cli, name, tmp_path)
    assert got == H.read_golden(name)


@pytest.mark.parametrize("name", ["A", "B", "D", "F", "I", "N", "O", "P"])
def test_clstr_identical_with_host_driven_phase_a_cpu(mock_cli, tmp_path, name):
    # MC_PHASE_A_STEPS=1: accumulate() driven from the host (host/bvec.hpp + one mc_accumulate_step per
    # scan) instead of mc_accumulate_run -- the path --align and unsupported histogram shapes take
    got, log = _run(mock_cli, name, tmp_path, env={"MC_PHASE_A_STEPS": "1"})
    assert "row compactions" in log
    assert got == H.read_golden(name)


@pytest.mark.parametrize("name", ["A", "B", "D", "E", "F"])
def test_clstr_identical_with_row_compaction_cpu(mock_cli, tmp_path, name):
    # rows that have joined a cluster are moved behind the alive ones as Phase A proceeds (here: from
    # 16 rows on, i.e. many times per run); the CLSTR file must not change
    got, log = _run(mock_cli, name, tmp_path, env={"MC_COMPACT_MIN_ROWS": "16", "MC_PHASE_A_STEPS": "1"})
    assert " 0 row compactions" not in log and "row compactions" in log
    assert got == H.read_golden(name)


@pytest.mark.parametrize("name", ["D", "F", "I"])
def test_clstr_identical_with_host_parser_cpu(mock_cli, tmp_path, name):
    # MC_HOST_PARSE=1: the host parser + mc_load_sequences instead of the index + mc_ingest_fasta the well-formed
    # inputs take by default (two files, IUPAC / lower case / N runs, 16-bit histograms): the same CLSTR file;
    # MC_SPLIT_FULL_SORT=1: Trainer::split's pivot sorts as full std::sorts instead of the lazy ones
    got, log = _run(mock_cli, name, tmp_path, env={"MC_HOST_PARSE": "1", "MC_SPLIT_FULL_SORT": "1"})
    assert "row order + segments" in log
    assert got == H.read_golden(name)
    got2, log2 = _run(mock_cli, name, tmp_path)
    assert "row order + segments" not in log2 and "row order" in log2
    assert got2 == got


@pytest.mark.parametrize("name,gpus", [("A", 2), ("N", 3)])
def test_clstr_identical_with_split_alignments_cpu(mock_cli, tmp_path, name, gpus):
    # the host side of `--gpus N` (batches of alignments cut by cells into one run per context, each driven by its own
    # host thread, sequences cloned once) against the CPU mock of the C-ABI: the CLSTR file must not change
    got, log = _run(mock_cli, name, tmp_path, extra=("--gpus", str(gpus)), env={"MC_ALIGN_SPLIT_MIN_CELLS": "1", "MC_ALIGN_SHARD_MIN_LEN": "1"})
    assert "sequences copied to %d more GPUs" % (gpus - 1) in log and "alignment batches split over the %d GPUs" % gpus in log
    assert got == H.read_golden(name), log[-1500:]


def test_gpus_flag_keeps_small_inputs_on_one_gpu_cpu(mock_cli, tmp_path):
    # --gpus N without anything to share (megabytes of input, records of a few hundred letters): the run says so, uses
    # one context and writes the same file
    got, log = _run(mock_cli, "A", tmp_path, extra=("--gpus", "4"))
    assert "--gpus 4:" in log and "one GPU" in log and "sequences copied" not in log and "peer inboxes connected" not in log
    assert got == H.read_golden("A")
    # --align shares the alignments whatever the record length
    got, log = _run(mock_cli, "G", tmp_path, extra=("--gpus", "2"))
    assert "alignments are split over the GPUs" in log
    assert got == H.read_golden("G")


def test_ingest_path_edge_records_cpu(mock_cli, tmp_path):
    # records the two input paths (index + mc_ingest_fasta / host parser + mc_load_sequences) must treat alike: all N and
    # one-letter records abort like the reference's segment->at(0) (Chromosome.cpp:193), records under 20 letters have
    # no segment but stay, an invalid letter is the InvalidInputException exit, N runs give several segments
    import random
    rnd = random.Random(3)
    base = ["".join(rnd.choice("ACGT") for _ in range(300)) for _ in range(40)]
    cases = {
        "allN": (base[:20] + ["N" * 50] + base[20:], False, "no usable sequence"),
        "one": (base[:5] + ["A"] + base[5:], False, "no usable sequence"),
        "short": (base[:5] + ["ACGTACGTAC"] + base[5:], True, ""),
        "bad": (base[:7] + [base[7][:100] + "!" + base[7][100:]] + base[8:], False, "Invalid nucleotide"),
        "Nrun": (base[:3] + [base[3][:100] + "N" * 30 + base[3][100:] + "nnnn"] + base[4:], True, ""),
    }
    for name, (recs, ok, msg) in cases.items():
        fa = os.path.join(str(tmp_path), name + ".fa")
        with open(fa, "w") as f:
            for i, sq in enumerate(recs):
                f.write(f">r{i} d\n" + "\n".join(sq[j:j + 60] for j in range(0, len(sq), 60)) + "\n")
        outs = []
        for env in ({}, {"MC_HOST_PARSE": "1"}):
            out = os.path.join(str(tmp_path), name + ".clstr")
            r = subprocess.run([mock_cli, fa, "--id", "0.9", "--kmer", "3", "--output", out], capture_output=True, text=True, env={**os.environ, **env})
            assert (r.returncode == 0) == ok and msg in r.stderr, (name, env, r.returncode, r.stderr[-300:])
            outs.append(open(out).read() if ok else r.stderr.strip().splitlines()[-1])
        assert outs[0] == outs[1], name


def test_cli_errors(mock_cli, tmp_path):
    # Runner.cpp:150-263: bad values and missing files exit non-zero with the reference's messages
    r = subprocess.run([mock_cli], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stdout
    r = subprocess.run([mock_cli, "--id", "1.5", __file__], capture_output=True, text=True)
    assert r.returncode == 1 and "Similarity must be between 0 and 1" in r.stderr
    r = subprocess.run([mock_cli, "/nonexistent.fa"], capture_output=True, text=True)
    assert r.returncode == 1
    r = subprocess.run([mock_cli, "--kmer", "0", __file__], capture_output=True, text=True)
    assert r.returncode == 1 and "K must be greater than 0" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(H.CASES))
def test_clstr_identical_to_reference_gpu(built_lib, tmp_path, name):
    from meshclust_b200 import build
    cli = build.build_cli()
    assert cli and os.path.exists(cli)
    got, log = _run(cli, name, tmp_path)
    assert got == H.read_golden(name), log[-1500:]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["A", "B", "D", "F", "I", "K", "N", "O", "P", "Q", "c1_full"])
def test_clstr_identical_with_host_driven_phase_a_gpu(built_lib, tmp_path, name):
    from meshclust_b200 import build
    cli = build.build_cli()
    got, log = _run(cli, name, tmp_path, env={"MC_PHASE_A_STEPS": "1"})
    assert "row compactions" in log
    assert got == H.read_golden(name), log[-1500:]


@pytest.mark.gpu
@pytest.mark.parametrize("name,gpus", [("B", 2), ("c1_full", 3), ("F", 2)])
def test_clstr_identical_with_sharded_phase_a(built_lib, tmp_path, name, gpus):
    # --gpus N shares the Phase-A scans between N contexts (on a 1-GPU box they share the device):
    # the CLSTR file must not change
    from meshclust_b200 import build
    cli = build.build_cli()
    got, log = _run(cli, name, tmp_path, extra=("--gpus", str(gpus)), env={"MC_PHASE_A_STEPS": "1"})
    assert "peer inboxes connected" in log
    assert got == H.read_golden(name), log[-1500:]


@pytest.mark.gpu
@pytest.mark.parametrize("name,gpus", [("A", 2), ("c1_full", 3), ("Q", 2), ("G", 2)])
def test_clstr_identical_with_split_alignments(built_lib, tmp_path, name, gpus):
    # --gpus N splits the alignment batches of the training stage by pairs over N contexts (on a
    # 1-GPU box they share the device), each with its own copy of the sequences; Phase A stays on one GPU
    from meshclust_b200 import build
    cli = build.build_cli()
    got, log = _run(cli, name, tmp_path, extra=("--gpus", str(gpus)), env={"MC_ALIGN_SPLIT_MIN_CELLS": "1", "MC_ALIGN_SHARD_MIN_LEN": "1"})
    assert "sequences copied to %d more GPUs" % (gpus - 1) in log and "peer inboxes connected" not in log
    assert got == H.read_golden(name), log[-1500:]


@pytest.mark.gpu
@pytest.mark.parametrize("name,gpus", [("A", 1), ("B", 1), ("c1_full", 1), ("E", 2), ("F", 3)])
def test_clstr_identical_with_row_compaction_gpu(built_lib, tmp_path, name, gpus):
    from meshclust_b200 import build
    cli = build.build_cli()
    got, log = _run(cli, name, tmp_path, extra=("--gpus", str(gpus)), env={"MC_COMPACT_MIN_ROWS": "16", "MC_PHASE_A_STEPS": "1"})
    assert " 0 row compactions" not in log and "row compactions" in log
    assert got == H.read_golden(name), log[-1500:]
