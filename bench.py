#!/usr/bin/env python
"""bench.py -- headline benchmark of the FM-index hot path (BASELINE.json configs[1]):

    exact search + locate of 10 M synthetic 150-bp reads against a synthetic 3 Gbp DNA BiFMIndex (rate 16)

  python bench.py --gpus N --steps K --warmup W              our arm  (libfmb200.so, hand-written sm_100a CUDA)
  python bench.py --impl reference --gpus N --steps K ...    reference arm: the reference's own CPU search + locate
                                                             (oracle/_ref/libfmref.so, all host threads), bounded sample

One "step" = one pass of search + locate over the whole read batch.  `value` = queries/s with the reads already
resident in HBM; `e2e` = the same through the C-ABI call fmb_search_and_locate with HOST buffers (pinned), H2D
and D2H inside the timed region.  Multi GPU (torchrun, one rank per GPU): the index is replicated per GPU, every
rank searches its own 10 M reads (weak scaling), no collective on the data path; the only collectives are the
barrier and the max-over-ranks of the device time.

Sizes can be reduced for smoke runs with --text / --reads / --read-len (the JSON line then names the reduced workload).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)"""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_workload(fmb, device, n_text, nq, L, rate, seed):
    """synthetic text on the device -> GPU index build -> reads copied from random text offsets (C2 of SURVEY §8d)"""
    from fmb200 import capi
    t0 = time.time()
    d_text = capi.synth_text_device(device, 5, n_text, seed)
    index = fmb.Index.build_from_device_text(5, d_text, n_text, sampling_rate=rate, bidirectional=True, device=device)
    t1 = time.time()
    d_reads = capi.synth_reads_device(device, d_text, n_text, nq, L, seed + 1)
    capi.device_free(device, d_text)
    sym = capi.PinnedArray(nq * L, np.uint8)
    capi.copy_to_host(device, sym.array, d_reads, nq * L)
    capi.device_free(device, d_reads)
    off = capi.PinnedArray(nq + 1, np.uint64)
    off.array[:] = np.arange(nq + 1, dtype=np.uint64) * np.uint64(L)
    log(f"index build {t1 - t0:.1f}s (n={n_text}), reads {time.time() - t1:.1f}s, device image {index.info.device_bytes / 1e9:.2f} GB")
    return index, sym, off


def reference_index(index):
    """host-side reference index over exactly the same BWT bytes and samples (BiFMIndex(bwt, bwtRev, SparseArray))"""
    from oracle.pyoracle import Ref
    t0 = time.time()
    bwt, rev, bm, sq, sp = index.export()
    t1 = time.time()
    ref = Ref.from_bwt(5, bwt, rev, bm, sq, sp)
    log(f"reference index: export {t1 - t0:.1f}s, construct {time.time() - t1:.1f}s")
    return ref


def cpu_search_locate(ref, sym, off, b, e, threads, L):
    """reference search_no_errors::search + LocateLinear on reads [b,e); returns (seconds, n_located)"""
    s = sym[b * L: e * L]
    o = (off[b: e + 1] - off[b]).astype(np.uint64)
    hits = ref.search_exact(s, o, threads=threads)
    t = ref.last_seconds
    locs = ref.locate(hits, threads=threads)
    t += ref.last_seconds
    return t, len(locs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--text", type=float, default=3e9, help="text length in symbols (default 3 Gbp)")
    ap.add_argument("--reads", type=float, default=1e7, help="reads per GPU (default 10 M)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--rate", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    n_text, nq, L = int(args.text), int(args.reads), args.read_len
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0

    import torch
    import fmb200 as fmb
    from fmb200 import capi

    if fmb.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libfmb200 has no CPU fallback")
    device = local_rank
    torch.cuda.set_device(device)
    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))

    workload = (f"exact search + locate, {nq} x {L}bp reads copied from the text, synthetic {n_text} bp DNA BiFMIndex "
                f"(sigma 5, sampling rate {args.rate}), per GPU")
    index, sym, off = build_workload(fmb, device, n_text, nq, L, args.rate, seed=3 + rank)
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        ref = reference_index(index)
        # size the per-step sample so that W+K steps take about 2 minutes in total
        probe = min(nq, 20000)
        t, _ = cpu_search_locate(ref, sym.array, off.array, 0, probe, threads, L)
        rate_qs = probe / max(t, 1e-9)
        per_step = int(min(nq, max(probe, rate_qs * 120.0 / (W + K))))
        for _ in range(W):
            cpu_search_locate(ref, sym.array, off.array, 0, per_step, threads, L)
        tot = 0.0
        for _ in range(K):
            t, nloc = cpu_search_locate(ref, sym.array, off.array, 0, per_step, threads, L)
            tot += t
        qps = per_step * K / tot
        sample = f"first {per_step} of the {nq} reads per step, search_no_errors::search (batched) + LocateLinear, {threads} threads"
        line = {"impl": "reference", "metric": "queries/s (150bp exact search + locate, 3 Gbp index)", "value": qps, "unit": "queries/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * tot / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": {"workload": workload},
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "reference", "sample": sample},
                "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------------------------------------------
    stream = torch.cuda.current_stream()
    capi.index_set_stream(index, stream.cuda_stream)
    queries = index.upload(sym.array, off.array)          # resident in HBM before the timed region

    def step():
        res = index.search_exact(queries)
        loc = index.locate(res)
        return res, loc

    # algorithmic work of the reference algorithm on this batch (occ-block lookups of one-symbol steps),
    # counted once, untimed, by the one-symbol kernel; the timed steps use the default (two-symbol + jump) kernel
    index.set_exact_mode(1)
    res = index.search_exact(queries)
    alg_lookups = res.stats.occ_lookups
    one_symbol_ms = res.stats.main_kernel_ms
    del res
    index.set_exact_mode(0)

    for _ in range(W):
        res, loc = step()
    n_hits, n_locs = len(res), len(loc)
    del res, loc

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(device)
    launches0 = capi.kernel_launch_count()
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    search_ms, locate_ms, lines_s, look_l = [], [], 0, 0
    for _ in range(K):
        res, loc = step()
        search_ms.append(res.stats.main_kernel_ms)
        locate_ms.append(loc.stats.main_kernel_ms)
        lines_s, look_l = res.stats.line_requests, loc.stats.lf_steps
        del res, loc
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = capi.kernel_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = nq * world / (ms_per_step * 1e-3)

    # ---- end to end through the C-ABI with host buffers ------------------------------------------------------
    out = capi.PinnedArray(nq + 1024, capi.LOC32_DTYPE)
    for _ in range(2):
        index.search_and_locate(sym.array, off.array, out=out.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        locs, _st = index.search_and_locate(sym.array, off.array, out=out.array)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_qps = nq * world * K / e2e_s
    h2d = nq * L + (nq + 1) * 8
    d2h = len(locs) * 16

    # ---- roofline of the dominant kernel (exact_search2_kernel) -----------------------------------------------------
    # achieved = ALGORITHMIC bytes (the reference algorithm's occ-block lookups x 32 B) / kernel time.
    # The kernel answers those lookups with far fewer physical line fetches (k-mer table, two-symbol lines, LF^16
    # jumps), so `frac` can exceed 1; `physical` is the kernel's own traffic (line requests x 128 B) against the same peak.
    peak, peak_src = load_peaks()
    k_ms = float(np.mean(search_ms))
    l_ms = float(np.mean(locate_ms))
    alg_bytes = alg_lookups * 32.0
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    phys_bytes = lines_s * 128.0
    loc_lookups = n_locs * 2 + look_l * 2            # per row: marker test + sample fetch, per LF step: occ block + marker word
    traffic = None                                   # dram bytes per launch from the committed ncu --set full capture, same workload only
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f)["exact_search2_kernel"]
        if ent["workload"] == f"{nq} x {L}bp on {n_text} bp":
            traffic = ent["traffic_bytes"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "exact_search2_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "lookups_per_query": alg_lookups / nq, "kernel_ms": k_ms,
                "physical": {"line_requests_per_query": lines_s / nq, "bytes_per_launch": phys_bytes,
                             "gbs": phys_bytes / (k_ms * 1e-3) / 1e9, "frac": phys_bytes / (k_ms * 1e-3) / 1e9 / peak,
                             "lines_per_s": lines_s / (k_ms * 1e-3), "measured_random_line_ceiling_per_s": 38.4e9},
                "one_symbol_kernel_ms": one_symbol_ms,
                "locate_kernel": {"kernel_ms": l_ms, "lf_steps_per_row": look_l / max(n_locs, 1), "algorithmic_bytes_per_launch": loc_lookups * 32.0,
                                  "achieved": loc_lookups * 32.0 / (l_ms * 1e-3) / 1e9, "frac": loc_lookups * 32.0 / (l_ms * 1e-3) / 1e9 / peak},
                "note": "frac > 1 is possible: algorithmic bytes are those of the reference's one-symbol algorithm; see physical.frac for the kernel's own traffic"}

    line = {"metric": "queries/s (150bp exact search + locate, 3 Gbp index)", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload, "l2_policy": "inputs larger than L2 (index image %.1f GB, reads %.2f GB)" % (index.info.device_bytes / 1e9, nq * L / 1e9),
                       "hits_per_step": n_hits, "located_rows_per_step": n_locs},
            "roofline": roofline,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks}

    # ---- CPU baseline: the reference's own search on this box's host cores (rank 0, N = 1 only) ---------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_index(index)
            probe = min(nq, 20000)
            t, _ = cpu_search_locate(ref, sym.array, off.array, 0, probe, threads, L)
            sample_n = int(min(nq, max(probe, probe / max(t, 1e-9) * args.cpu_seconds)))
            t, nloc = cpu_search_locate(ref, sym.array, off.array, 0, sample_n, threads, L)
            t1, _ = cpu_search_locate(ref, sym.array, off.array, 0, min(sample_n, 50000), 1, L)
            line["cpu_baseline"] = {"value": sample_n / t, "unit": "queries/s", "cores": threads, "kind": "reference",
                                    "sample": f"first {sample_n} of the {nq} reads, search_no_errors::search + LocateLinear, {threads} threads "
                                              f"({t:.1f}s); 1 thread: {min(sample_n, 50000) / t1:.0f} queries/s"}
            # parity spot check on the sample: located rows identical as sorted sets
            got = np.sort(locs[locs["qidx"] < sample_n], order=["qidx", "seq", "pos", "e"])
            hits = ref.search_exact(sym.array[: sample_n * L], off.array[: sample_n + 1], threads=threads)
            exp = ref.locate(hits, threads=threads)
            exp32 = np.zeros(len(exp), dtype=capi.LOC32_DTYPE)
            for f in ("qidx", "seq", "pos", "e"):
                exp32[f] = exp[f]
            exp32 = np.sort(exp32, order=["qidx", "seq", "pos", "e"])
            line["parity"] = {"checked_queries": sample_n, "identical": bool(np.array_equal(got, exp32))}
        except Exception as ex:   # the baseline must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": threads, "kind": "reference", "sample": f"failed: {ex}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
