"""Write a BASELINE config as FASTA: python tools/gen_config.py c2 /tmp/c2.fa [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshclust_b200 import synth
name, path = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else None
l, o, t = synth.generate_config(name, n)
synth.write_fasta(path, l, o, synth.headers_for(o.size - 1, t))
print(name, o.size - 1, "sequences", int(o[-1]), "bases ->", path)
