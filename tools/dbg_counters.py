import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fmb200 as fmb
from fmb200 import capi, schemes
from oracle.pyoracle import Oracle, Counters, sort_hits
n_text, nq, L = int(1e8), 20000, 150
d_text = capi.synth_text_device(0, 5, n_text, 3)
index = fmb.Index.build_from_device_text(5, d_text, n_text, sampling_rate=16, bidirectional=True, device=0)
bwt, rev, bm, sq, sp = index.export()
o = Oracle.from_bwt(5, bwt, rev, bm, sq, sp)
off = np.arange(nq + 1, dtype=np.uint64) * np.uint64(L)
for k, edit in ((1, False), (2, True)):
    d_reads = capi.synth_reads_err_device(0, d_text, n_text, nq, L, 5, 5, k, edit)
    sym = np.zeros(nq * L, dtype=np.uint8)
    capi.copy_to_host(0, sym, d_reads, nq * L)
    sch = schemes.optimum(0, k)
    part = schemes.uniform_partition(sch[0].shape[1], L)
    res = index.search_scheme(index.upload(sym, off), sch, part, edit)
    ctr = Counters()
    exp = o.search_ng26(sym, off, sch, part, edit, counters=ctr)
    same = np.array_equal(sort_hits(res.hits()), sort_hits(exp))
    print(f"k={k} edit={edit}: hits equal {same}; gpu ext {res.stats.extensions} oracle ext {ctr.extensions}; gpu look {res.stats.occ_lookups} oracle look {ctr.occ_lookups}; phys {res.stats.line_requests}")
