"""where does the end-to-end time of fmb_search_and_locate go?  (FMB_TRACE=1; run on the GPU box)
   python tools/e2e_trace.py [workload] [reads] [text]"""
import os, sys, time
os.environ.setdefault("FMB_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fmb200 as fmb
from fmb200 import capi
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "k1-edit"
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
n_text = int(float(sys.argv[3])) if len(sys.argv) > 3 else 3_000_000_000
L = 150
d_text = capi.synth_text_device(0, 5, n_text, 3)
index = fmb.Index.build_from_device_text(5, d_text, n_text, sampling_rate=16, bidirectional=True, device=0)
sym, off = bench.make_reads(capi, 0, wl, d_text, n_text, nq, L, 3, 4, 5)
capi.device_free(0, d_text)
scheme, partition, edit, k = bench.scheme_of(wl, L)
out = capi.PinnedArray(nq * 5 + 1024, capi.LOC32_DTYPE)
packed = capi.pack_queries(sym.array, 5)
pw = capi.PinnedArray(packed[0].size, np.uint32)
pw.array[:] = packed[0]
for name, kw in (("bytes", dict(symbols=sym.array)), ("packed", dict(symbols=None, packed=(pw.array, packed[1], packed[2])))):
    for rep in range(4):
        print(f"---- {wl} {name} rep {rep}", file=sys.stderr, flush=True)
        t1 = time.perf_counter()
        locs, st = index.search_and_locate(kw["symbols"], off.array, scheme=scheme, partition=partition, edit=edit, out=out.array, packed=kw.get("packed"))
        ms = 1e3 * (time.perf_counter() - t1)
        print(f"  {wl} {name}: call {ms:.1f} ms = {nq / ms / 1e3:.1f} M q/s, rows {len(locs)}, sum kernel_ms {st.kernel_ms:.1f}", file=sys.stderr, flush=True)
