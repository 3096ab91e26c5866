"""Generates tests/golden/ref_vectors.json by running the REFERENCE ITSELF (oracle/_ref/libfmref.so, built from
/root/reference by oracle/build_ref.sh) on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The JSON travels with the repo; tests compare the C oracle (CPU) and libfmb200 (GPU) against it, so parity is
anchored on outputs of the reference even where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import fmb200  # noqa: E402
from fmb200 import schemes, synth  # noqa: E402
from oracle.pyoracle import Ref, build_ref, sort_hits, sort_locs  # noqa: E402


def hits_list(h):
    h = sort_hits(h)
    return [[int(x) for x in r] for r in h.tolist()]


def locs_list(l):
    l = sort_locs(l)
    return [[int(x) for x in r] for r in l.tolist()]


def scheme_list(s):
    return [a.astype(int).tolist() for a in s]


def hits_in_order(h):
    """hit-limited searches: the delegate calls in the order the reference makes them (one thread: ascending qidx, depth-first
    discovery order inside a query) -- NOT sorted"""
    return [[int(x) for x in r] for r in h.tolist()]


def repeat_text(seed):
    """30 mutated copies of a 40-symbol unit, one sequence each: wide intervals, many rows per cursor"""
    rng = np.random.default_rng(seed)
    unit = rng.integers(1, 5, 40).astype(np.uint8)
    seqs = []
    for _ in range(30):
        u = unit.copy()
        u[rng.integers(0, 40, 2)] = rng.integers(1, 5, 2)
        seqs.append(u)
    return np.concatenate([np.concatenate([s, [0]]) for s in seqs]).astype(np.uint8), seqs


def main():
    assert build_ref(), "reference library not available"
    out = {"schemes": {}, "cases": []}
    # scheme tables straight from the reference's generator registry (search_scheme/generator/all.h)
    for name in ("optimum", "pigeon_opt", "pigeon", "backtracking", "h2-k1", "h2-k2", "h2-k3", "kucherov-k1", "kucherov-k2", "01*0", "suffixFilter"):
        for (mn, mx) in ((0, 0), (0, 1), (0, 2), (1, 2), (0, 3)):
            try:
                out["schemes"][f"{name}:{mn}:{mx}"] = scheme_list(Ref.scheme(name, mn, mx))
            except Exception:
                pass
    cases = [
        dict(name="multi_seq_rate4", lengths=[1500, 700, 90, 1], rate=4, seed=11, L=24, nq=40),
        dict(name="single_seq_rate16", lengths=[4000], rate=16, seed=12, L=30, nq=40),
        dict(name="tiny_rate1", lengths=[40, 33], rate=1, seed=13, L=8, nq=30),
    ]
    for c in cases:
        text = synth.multi_text(c["lengths"], 5, c["seed"])
        ref = Ref.build(text, 5, c["rate"], True)
        n = ref.n
        bwt = [int(ref.symbol(i, 0)) for i in range(n)]
        bwt_rev = [int(ref.symbol(i, 1)) for i in range(n)]
        locate_all = [list(ref.locate_row(i)) for i in range(n)]
        first = c["lengths"][0]
        reads, _ = synth.reads_from_text(text[: first + 1], c["nq"], c["L"], c["seed"] + 1)
        reads_e = synth.plant_errors(reads, 5, 1, True, c["seed"] + 2)
        reads_all = np.concatenate([reads[: c["nq"] // 2], reads_e[c["nq"] // 2:]])
        sym, off = synth.flatten(reads_all)
        entry = dict(c)
        entry.update(text=text.tolist(), bwt=bwt, bwt_rev=bwt_rev, C=[int(x) for x in ref.C], locate_all=locate_all,
                     queries=reads_all.tolist(), searches={})
        h = ref.search_exact(sym, off)
        entry["searches"]["exact"] = dict(hits=hits_list(h), locs=locs_list(ref.locate(h)))
        for k in (1, 2):
            for edit in (False, True):
                sch = schemes.optimum(0, k)
                part = schemes.uniform_partition(sch[0].shape[1], c["L"])
                h = ref.search_ng26(sym, off, sch, part, edit)
                entry["searches"][f"ng26_optimum_k{k}_{'edit' if edit else 'ham'}"] = dict(hits=hits_list(h), locs=locs_list(ref.locate(h)))
                h = ref.search_facade(sym, off, edit, k)
                entry["searches"][f"facade_k{k}_{'edit' if edit else 'ham'}"] = dict(hits=hits_list(h))
                for lim in (1, 3):     # search_ng26::search(..., n): SearchNg26.h:408-433
                    h = ref.search_ng26(sym, off, sch, part, edit, max_hits=lim)
                    entry["searches"][f"ng26_optimum_k{k}_{'edit' if edit else 'ham'}_n{lim}"] = dict(hits_in_order=hits_in_order(h))
        short = reads_all[:, :10]
        ssym, soff = synth.flatten(short)
        entry["short_queries"] = short.tolist()
        for k in (0, 1, 2):
            h = ref.search_backtracking(ssym, soff, k)
            entry["searches"][f"backtracking_k{k}"] = dict(hits=hits_list(h))
        # String_c samples
        rng = np.random.default_rng(c["seed"])
        rows = sorted(set(int(x) for x in rng.integers(0, n + 1, size=60)) | {0, n})
        entry["string"] = [dict(row=r, dir=d, rank=[int(ref.rank(r, s, d)) for s in range(5)],
                                prefix_rank=[int(ref.prefix_rank(r, s, d)) for s in range(6)]) for r in rows for d in (0, 1)]
        out["cases"].append(entry)
    # hit limit on a repetitive collection: the order of the children of wide nodes and the clipping of cursors with many rows
    text, seqs = repeat_text(21)
    ref = Ref.build(text, 5, 2, True)
    reads = np.array([seqs[i % 30][4:20] for i in range(40)], dtype=np.uint8)
    reads = synth.plant_errors(reads, 5, 1, True, 22)
    sym, off = synth.flatten(reads)
    entry = dict(name="repeats_hit_limit", rate=2, L=16, text=text.tolist(), queries=reads.tolist(), searches={})
    for k in (1, 2):
        for edit in (False, True):
            sch = schemes.optimum(0, k)
            part = schemes.uniform_partition(sch[0].shape[1], 16)
            for lim in (1, 4, 25, 1000):
                h = ref.search_ng26(sym, off, sch, part, edit, max_hits=lim)
                entry["searches"][f"ng26_optimum_k{k}_{'edit' if edit else 'ham'}_n{lim}"] = dict(hits_in_order=hits_in_order(h))
    out["hit_limit_case"] = entry
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
