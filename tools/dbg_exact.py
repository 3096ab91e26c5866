import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fmb200 as fmb
from fmb200 import synth
from helpers import make_index_pair
from oracle.pyoracle import sort_hits
text = synth.multi_text([40000], 5, 21)
o, g = make_index_pair(fmb, text, 5, 4)
g.set_exact_mode(2)
rng = np.random.default_rng(5)
reads = []
for L in (1, 2, 3, 4, 7, 8, 20, 21, 64):
    for _ in range(60):
        p = int(rng.integers(0, text.size - L))
        reads.append(text[p:p + L].copy())
    for _ in range(10):
        reads.append(rng.integers(0, 5, size=L).astype(np.uint8))
for a in range(5):
    for b in range(5):
        reads.append(np.array([a, b], dtype=np.uint8))
        for c in range(5):
            reads.append(np.array([a, b, c], dtype=np.uint8))
sym, off = synth.flatten(reads)
res = g.search_exact(g.upload(sym, off))
got = {int(h["qidx"]): (int(h["lb"]), int(h["len"])) for h in res.hits()}
exp = {int(h["qidx"]): (int(h["lb"]), int(h["len"])) for h in o.search_exact(sym, off)}
bad = 0
for q in range(len(reads)):
    if got.get(q) != exp.get(q):
        bad += 1
        if bad < 15:
            print("q", q, "read", reads[q].tolist(), "got", got.get(q), "exp", exp.get(q))
print("bad", bad, "of", len(reads), "kmer/jump info", g.info.device_bytes)
