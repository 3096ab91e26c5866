"""per-launch times of a k-error search (FMB_TRACE_SCHEME=1): python tools/scheme_trace.py [text] [reads] [k] [edit] [sigma] [read length]"""
import os, sys, time
os.environ["FMB_TRACE_SCHEME"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fmb200 as fmb
from fmb200 import capi, schemes
n_text = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
ks = [int(sys.argv[3])] if len(sys.argv) > 3 else [1, 2]
edits = [bool(int(sys.argv[4]))] if len(sys.argv) > 4 else [False, True]
sigma = int(sys.argv[5]) if len(sys.argv) > 5 else 5
L = int(sys.argv[6]) if len(sys.argv) > 6 else 150
d_text = capi.synth_text_device(0, sigma, n_text, 3)
index = fmb.Index.build_from_device_text(sigma, d_text, n_text, sampling_rate=16, bidirectional=True, device=0)
off = np.arange(nq + 1, dtype=np.uint64) * np.uint64(L)
for k in ks:
    for edit in edits:
        d_reads = capi.synth_reads_err_device(0, d_text, n_text, nq, L, 5, sigma, k, edit)
        sym = np.zeros(nq * L, dtype=np.uint8)
        capi.copy_to_host(0, sym, d_reads, nq * L)
        capi.device_free(0, d_reads)
        q = index.upload(sym, off)
        sch = schemes.optimum(0, k)
        part = schemes.uniform_partition(sch[0].shape[1], L)
        for rep in range(2):
            print(f"---- k={k} {'edit' if edit else 'hamming'} rep {rep}", file=sys.stderr, flush=True)
            res = index.search_scheme(q, sch, part, edit)
        st = res.stats
        print(f"k={k} {'edit' if edit else 'hamming'}: kernel {st.main_kernel_ms:.2f} ms  {nq / st.main_kernel_ms / 1e3:.2f} M q/s  hits {len(res)}  "
              f"extensions/q {st.extensions / nq:.1f}  requests/q {st.line_requests / nq:.1f}", flush=True)
        del q, res
