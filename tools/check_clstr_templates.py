"""Full-size sanity of a CLSTR file written for a synthetic BASELINE config (tools/gen_config.py headers carry the
template a sequence was mutated from): every sequence appears exactly once, and how clusters and templates relate.
python tools/check_clstr_templates.py out.clstr n_expected"""
import collections, re, sys
path, n_expected = sys.argv[1], int(sys.argv[2])
clusters, cur, seen, stars = [], None, set(), 0
for line in open(path):
    if line.startswith(">Cluster"):
        cur = []
        clusters.append(cur)
        continue
    m = re.search(r">seq(\d+) template(\d+)", line)
    assert m, line
    sid, t = int(m.group(1)), int(m.group(2))
    assert sid not in seen, f"sequence {sid} appears twice"
    seen.add(sid)
    cur.append(t)
    stars += line.rstrip().endswith("*")
assert len(seen) == n_expected, (len(seen), n_expected)
pure = sum(1 for c in clusters if len(set(c)) == 1)
by_t = collections.defaultdict(set)
for i, c in enumerate(clusters):
    for t in c:
        by_t[t].add(i)
whole = sum(1 for t, cs in by_t.items() if len(cs) == 1)
sizes = sorted(len(c) for c in clusters)
print(f"{path}: {len(seen)} sequences, each exactly once; {len(clusters)} clusters ({stars} with a '*' member), sizes {sizes[0]}..{sizes[-1]}; "
      f"{pure} clusters hold one template only; {whole} of {len(by_t)} templates lie in one cluster")
