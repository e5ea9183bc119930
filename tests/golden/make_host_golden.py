"""Generate tests/golden/clstr/*.clstr.gz with the COMPILED, UNMODIFIED reference CLI
(oracle/_ref/meshclust --threads 1, the deterministic parity oracle of SURVEY.md section 0).
Build container only:   python tests/golden/make_host_golden.py [case ...]"""
import gzip
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _hostcases as H  # noqa: E402
import _oracle as O  # noqa: E402

os.makedirs(H.GOLDEN_DIR, exist_ok=True)
for name in (sys.argv[1:] or list(H.CASES)):
    with tempfile.TemporaryDirectory() as d:
        paths, args = H.make_inputs(name, d)
        out = os.path.join(d, "ref.clstr")
        subprocess.run([O.REF_BIN, *paths, *args, "--threads", "1", "--output", out], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        data = open(out, "rb").read()
        with gzip.GzipFile(H.golden_path(name), "wb", mtime=0) as f:
            f.write(data)
        print(name, len(data), "bytes,", data.count(b">Cluster"), "clusters")
