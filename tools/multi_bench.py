"""One process driving N GPUs through ONE call: fmb_index_replicate + fmb_search_and_locate_multi.
   python tools/multi_bench.py [n_gpus] [reads_total] [text] [workloads...]       (run with gpurun --gpus N)
Prints one JSON line per workload: replicate time, end-to-end queries/s of the sharded call (host buffers in, rows out), parity of
the rows against the single-GPU call."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fmb200 as fmb
from fmb200 import capi
import bench
G = int(sys.argv[1]) if len(sys.argv) > 1 else fmb.device_count()
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
n_text = int(float(sys.argv[3])) if len(sys.argv) > 3 else 3_000_000_000
wls = sys.argv[4:] or ["exact", "k1-edit"]
L = 150
os.environ.setdefault("FMB_E2E_THREADS", str(max(2, min(6, (os.cpu_count() or 6) // G))))
t0 = time.time()
d_text = capi.synth_text_device(0, 5, n_text, 3)
index = fmb.Index.build_from_device_text(5, d_text, n_text, sampling_rate=16, bidirectional=True, device=0)
t1 = time.time()
data = {wl: bench.make_reads(capi, 0, wl, d_text, n_text, nq, L, 3, 4 + 17 * i, 5) for i, wl in enumerate(wls)}
capi.device_free(0, d_text)
replicas = [index]
t2 = time.time()
for g in range(1, G):
    replicas.append(index.replicate(g))
t3 = time.time()
print(f"[multi] build {t1 - t0:.1f}s, {G - 1} replica(s) of a {index.info.device_bytes / 1e9:.1f} GB image in {t3 - t2:.2f}s "
      f"({(G - 1) * index.info.device_bytes / 1e9 / max(t3 - t2, 1e-9):.0f} GB/s)", file=sys.stderr, flush=True)
for wl in wls:
    sym, off = data[wl]
    scheme, partition, edit, k = bench.scheme_of(wl, L)
    cap = nq * (6 if edit else 2) // G + 1024
    out = capi.PinnedArray(cap * G, capi.LOC32_DTYPE)
    single, _ = index.search_and_locate(sym.array, off.array, scheme=scheme, partition=partition, edit=edit, capacity=nq * 6 if edit else nq * 2)
    single = np.sort(single, order=["qidx", "seq", "pos", "e"])
    times = []
    for rep in range(5):
        ta = time.perf_counter()
        parts, st = capi.search_and_locate_multi(replicas, sym.array, off.array, scheme=scheme, partition=partition, edit=edit, shard_capacity=cap, out=out.array)
        times.append(time.perf_counter() - ta)
    rows = np.sort(np.concatenate(parts), order=["qidx", "seq", "pos", "e"])
    t1g = []
    for rep in range(3):
        ta = time.perf_counter()
        index.search_and_locate(sym.array, off.array, scheme=scheme, partition=partition, edit=edit, out=out.array)
        t1g.append(time.perf_counter() - ta)
    best = min(times[1:])
    print(json.dumps({"workload": wl, "n_gpus": G, "reads_total": nq, "scaling": "strong", "mode": "one process, fmb_search_and_locate_multi",
                      "e2e_queries_per_s": nq / best, "ms_per_call": 1e3 * best, "single_gpu_e2e_queries_per_s": nq / min(t1g[1:]),
                      "speedup_over_one_gpu": min(t1g[1:]) / best, "rows": int(rows.size), "rows_identical_to_single_gpu_call": bool(np.array_equal(rows, single)),
                      "replicate_seconds": t3 - t2, "image_gb": index.info.device_bytes / 1e9}), flush=True)
