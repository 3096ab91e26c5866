"""examples/fmb_example.cpp end to end: FASTA reference + FASTA reads -> index (cached on disk) -> k-error
search of the reads and their reverse complements -> "queryId seqId pos" lines, compared with the oracle on the same data."""
import os
import subprocess

import numpy as np
import pytest

from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMP = np.array([0, 4, 3, 2, 1], dtype=np.uint8)


def _fasta(path, seqs, prefix):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            body = "".join("$ACGT"[c] for c in s)
            f.write(f">{prefix}{i}\n")
            for j in range(0, len(body), 60):
                f.write(body[j:j + 60] + "\n")


def _lines(path):
    return sorted(tuple(int(x) for x in line.split()) for line in open(path))


def test_example_program(gpu, tmp_path):
    from fmb200 import schemes, synth
    lib = os.path.join(ROOT, "fmindex-collection_b200")
    exe = str(tmp_path / "fmb_example")
    subprocess.run(["g++", "-std=c++20", "-O1", "-Wno-comment", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "fmb_example.cpp"),
                    "-L", lib, "-lfmb200", f"-Wl,-rpath,{lib}", "-o", exe], check=True)
    rng = np.random.default_rng(8)
    seqs = [rng.integers(1, 5, 7000).astype(np.uint8), rng.integers(1, 5, 2500).astype(np.uint8)]
    text = np.concatenate([np.concatenate([s, [0]]) for s in seqs]).astype(np.uint8)
    L = 36
    reads = []
    for i in range(150):
        s = seqs[i % 2]
        p = int(rng.integers(0, len(s) - L))
        r = s[p:p + L].copy()
        if i % 2:
            r = COMP[r[::-1]]
        reads.append(r)
    reads = list(synth.plant_errors(np.array(reads, dtype=np.uint8), 5, 1, True, 3))
    ref_fa, reads_fa, out = str(tmp_path / "ref.fa"), str(tmp_path / "reads.fa"), str(tmp_path / "out.txt")
    _fasta(ref_fa, seqs, "chr")
    _fasta(reads_fa, reads, "read")
    doubled = []
    for r in reads:
        doubled += [r, COMP[r[::-1]]]
    sym, off = synth.flatten(doubled)
    o = Oracle.build(text, 5, 16)

    def expected(hits):
        return sorted((int(x["qidx"]), int(x["seq"]), int(x["pos"])) for x in o.locate(hits))

    # --mode all, edit distance k = 1 (fmc::Search -> fmc::search<true>, h2 scheme)
    r = subprocess.run([exe, "--ref", ref_fa, "--query", reads_fa, "--max_k", "1", "--save_output", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(ref_fa + ".fmb")
    sch, part = schemes.facade_scheme(True, 1, L)
    exp = expected(o.search_ng26(sym, off, sch, part, True))
    assert _lines(out) == exp and len(exp) >= 100
    # the same search on 2-bit packed reads (io.hpp packs while parsing, fmb_search_and_locate_packed): identical lines
    r = subprocess.run([exe, "--ref", ref_fa, "--query", reads_fa, "--max_k", "1", "--packed", "--save_output", out + "p"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "(packed)" in r.stdout and _lines(out + "p") == exp
    # second run: index loaded from the cache file; Hamming, hit limit 1 (search_n)
    r = subprocess.run([exe, "--ref", ref_fa, "--query", reads_fa, "--max_k", "2", "--hamming", "--maxhitsperquery", "1", "--save_output", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    sch, part = schemes.facade_scheme(False, 2, L)
    assert _lines(out) == expected(o.search_ng26(sym, off, sch, part, False, max_hits=1))
    # besthits: per query the lowest error level with a hit (search_best, SearchNg26.h:448-470), at most 2 rows per level
    r = subprocess.run([exe, "--ref", ref_fa, "--query", reads_fa, "--max_k", "2", "--mode", "besthits", "--maxhitsperquery", "2", "--save_output", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    best, done = [], set()
    for k in range(3):
        sch, part = schemes.facade_scheme(True, k, L)
        h = o.search_ng26(sym, off, sch, part, True, max_hits=2)
        keep = h[~np.isin(h["qidx"], list(done))] if done else h
        best += expected(keep)
        done |= set(int(q) for q in h["qidx"])
    assert _lines(out) == sorted(best)
    # without reverse complements only the forward-strand reads are found
    r = subprocess.run([exe, "--ref", ref_fa, "--query", reads_fa, "--max_k", "1", "--no-reverse", "--save_output", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    fsym, foff = synth.flatten(reads)
    sch, part = schemes.facade_scheme(True, 1, L)
    assert _lines(out) == expected(o.search_ng26(fsym, foff, sch, part, True))
