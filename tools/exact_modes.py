"""one-symbol vs two-symbol exact search kernel on the bench workload (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fmb200 as fmb
import bench

n_text = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
L = int(sys.argv[3]) if len(sys.argv) > 3 else 150
index, sym, off = bench.build_workload(fmb, 0, n_text, nq, L, 16, 3)
q = index.upload(sym.array, off.array)
ref = None
for mode in (1, 2, 1, 2):
    index.set_exact_mode(mode)
    ms = []
    for _ in range(4):
        res = index.search_exact(q)
        ms.append(res.stats.main_kernel_ms)
    st = res.stats
    h = res.hits()
    if ref is None:
        ref = h
    same = np.array_equal(ref, h)
    print(f"mode {mode}: kernel {np.mean(ms[1:]):.2f} ms  ({nq / np.mean(ms[1:]) / 1e3:.1f} M q/s)  hits {len(h)} identical {same}  "
          f"lookups/q {st.occ_lookups / nq:.2f}  line_requests/q {st.line_requests / nq:.2f}", flush=True)
