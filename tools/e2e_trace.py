"""where does the end-to-end time go?  (run on the GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fmb200 as fmb
from fmb200 import capi
import bench

n_text, nq, L = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000, int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000, 150
index, sym, off = bench.build_workload(fmb, 0, n_text, nq, L, 16, 3)
def T(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
# raw pinned H2D with torch
pin = torch.from_numpy(sym.array)
dev = torch.empty(nq * L, dtype=torch.uint8, device="cuda")
print("torch H2D 1.5GB (numpy view of pinned memory): %.2f ms" % T(lambda: dev.copy_(pin, non_blocking=True)))
print("fmb_queries_upload: %.2f ms" % T(lambda: index.upload(sym.array, off.array)))
q = index.upload(sym.array, off.array)
print("search_exact: %.2f ms" % T(lambda: index.search_exact(q)))
res = index.search_exact(q)
print("locate: %.2f ms" % T(lambda: index.locate(res)))
loc = index.locate(res)
print("fetch locs32 (pageable numpy): %.2f ms" % T(lambda: loc.locs32()))
out = capi.PinnedArray(nq + 1024, capi.LOC32_DTYPE)
print("search_and_locate: %.2f ms" % T(lambda: index.search_and_locate(sym.array, off.array, out=out.array)))
