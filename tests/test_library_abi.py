"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/meshclust_b200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "meshclust_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(built_lib.EXPORTS)


def test_sass_is_sm100a(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_host_segments_matches_oracle(built_lib, oracle, golden):
    offs, letters, seg_off = golden["enc_offs"], golden["enc_letters"], golden["enc_seg_off"]
    for i in range(offs.size - 1):
        s = built_lib.host_segments(letters[offs[i]:offs[i + 1]])
        assert np.array_equal(s.reshape(-1), golden["enc_segs"][2 * seg_off[i]:2 * seg_off[i + 1]])
    assert built_lib.host_segments(b"NNNNNN") is None
    assert built_lib.host_segments(b"NNNNNA") is None          # run starting on the last char is dropped
    assert len(built_lib.host_segments(b"ACGTACGTAC")) == 0    # < 20 bp: no segment, but valid
    rng = np.random.default_rng(3)
    alpha = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)
    for _ in range(300):
        L = int(rng.integers(1, 200))
        s = rng.choice(alpha[:8], L)
        for _ in range(int(rng.integers(0, 5))):
            st, ln = int(rng.integers(0, L)), int(rng.integers(1, 14))
            s[st:st + ln] = ord("N") if rng.random() < 0.5 else ord("n")
        d, segs = oracle.encode(s.tobytes())
        mine = built_lib.host_segments(s)
        if d is None:
            assert mine is None
        else:
            assert np.array_equal(mine, segs)


def test_no_gpu_is_a_loud_error(built_lib):
    if built_lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(built_lib.McError) as e:
        built_lib.Context(0)
    assert e.value.code == built_lib.MC_ERR_CUDA
    assert "no CPU fallback" in str(e.value)
