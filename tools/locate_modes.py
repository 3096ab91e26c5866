"""locate kernel variants on the bench workload (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fmb200 as fmb
import bench
n_text = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
index, sym, off = bench.build_workload(fmb, 0, n_text, nq, 150, 16, 3)
q = index.upload(sym.array, off.array)
res = index.search_exact(q)
ms = []
for _ in range(4):
    loc = index.locate(res)
    ms.append(loc.stats.main_kernel_ms)
print(f"locate kernel {np.mean(ms[1:]):.3f} ms, rows {len(loc)}, lf steps/row {loc.stats.lf_steps / len(loc):.2f}, image {index.info.device_bytes / 1e9:.2f} GB")
