"""end-to-end (host buffers) time of a k-error workload for the chunking experiments"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fmb200 as fmb
from fmb200 import capi
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "k1-hamming"
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
index, sym, off = bench.build_workload(fmb, 0, 3_000_000_000, nq, 150, 16, 3, workload=wl)
scheme, partition, edit = bench.workload_scheme(wl, 150)
out = capi.PinnedArray(nq * 5 + 1024, capi.LOC32_DTYPE)
for _ in range(2):
    index.search_and_locate(sym.array, off.array, scheme=scheme, partition=partition, edit=edit, out=out.array)
t0 = time.perf_counter()
for _ in range(5):
    t1 = time.perf_counter()
    locs, st = index.search_and_locate(sym.array, off.array, scheme=scheme, partition=partition, edit=edit, out=out.array)
    print(f"  call {1e3 * (time.perf_counter() - t1):.1f} ms", flush=True)
t2 = time.perf_counter()
torch.cuda.synchronize()
print(f"  final synchronize {1e3 * (time.perf_counter() - t2):.1f} ms; loop {1e3 * (t2 - t0):.1f} ms", flush=True)
ms = (t2 - t0) / 5 * 1e3
print(f"{wl}: e2e {ms:.1f} ms  {nq / ms / 1e3:.1f} M q/s  rows {len(locs)}  sum kernel_ms {st.kernel_ms:.1f}")
