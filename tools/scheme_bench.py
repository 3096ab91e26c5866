"""k-error scheme search at scale (run on the GPU box): kernel time, work counters, spot parity against the reference"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fmb200 as fmb
from fmb200 import capi, schemes
n_text = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
L = 150
t0 = time.time()
d_text = capi.synth_text_device(0, 5, n_text, 3)
index = fmb.Index.build_from_device_text(5, d_text, n_text, sampling_rate=16, bidirectional=True, device=0)
print(f"index build {time.time() - t0:.1f}s image {index.info.device_bytes / 1e9:.1f} GB", flush=True)
off = np.arange(nq + 1, dtype=np.uint64) * np.uint64(L)
for k in (1, 2):
    for edit in (False, True):
        d_reads = capi.synth_reads_err_device(0, d_text, n_text, nq, L, 5, 5, k, edit)
        sym = np.zeros(nq * L, dtype=np.uint8)
        capi.copy_to_host(0, sym, d_reads, nq * L)
        capi.device_free(0, d_reads)
        q = index.upload(sym, off)
        sch = schemes.optimum(0, k)
        part = schemes.uniform_partition(sch[0].shape[1], L)
        ms = []
        for _ in range(3):
            res = index.search_scheme(q, sch, part, edit)
            ms.append(res.stats.main_kernel_ms)
        st = res.stats
        print(f"k={k} {'edit' if edit else 'hamming'}: kernel {np.mean(ms[1:]):.2f} ms  {nq / np.mean(ms[1:]) / 1e3:.2f} M q/s  hits {len(res)}  "
              f"extensions/q {st.extensions / nq:.1f}  lookups/q {st.occ_lookups / nq:.1f}  frontier_peak {st.frontier_peak}", flush=True)
