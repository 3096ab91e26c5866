#!/usr/bin/env python
"""bench.py -- headline benchmark of the FM-index hot path (BASELINE.json configs[1]):

    exact search + locate of 10 M synthetic 150-bp reads against a synthetic 3 Gbp DNA BiFMIndex (rate 16)

  python bench.py --gpus N --steps K --warmup W              our arm  (libfmb200.so, hand-written sm_100a CUDA)
  python bench.py --workload k2-edit ...                     the other BASELINE configs (k-error schemes, locate-heavy)
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


WORKLOADS = ("exact", "k1-hamming", "k1-edit", "k2-hamming", "k2-edit", "locate-heavy", "protein-k1-hamming", "protein-k1-edit")


def build_workload(fmb, device, n_text, nq, L, rate, seed, workload="exact", read_seed=None):
    """synthetic text on the device -> GPU index build -> reads:
         exact          reads copied from random text offsets (C2)
         kK-hamming/edit  the same reads with e in {0..K} planted substitutions / substitutions+indels each (C3)
         locate-heavy   text with a repeat family (1000 copies of a 1 kbp unit, 1 % substitutions), reads = windows of the unit (C4)"""
    from fmb200 import capi
    t0 = time.time()
    read_seed = seed + 1 if read_seed is None else read_seed
    sigma = 21 if workload.startswith("protein") else 5
    if workload == "locate-heavy":
        d_text = capi.synth_repeat_text_device(device, 5, n_text, seed, 1000, 1000, 10)
    else:
        d_text = capi.synth_text_device(device, sigma, n_text, seed)
    index = fmb.Index.build_from_device_text(sigma, d_text, n_text, sampling_rate=rate, bidirectional=True, device=device)
    t1 = time.time()
    if workload == "exact":
        d_reads = capi.synth_reads_device(device, d_text, n_text, nq, L, read_seed)
    elif workload == "locate-heavy":
        d_reads = capi.synth_unit_reads_device(device, 5, nq, L, seed, 1000)   # unit = f(seed); reads differ per rank through the offset hash below
        # (the unit-window hash takes the read seed through nq-independent q; ranks use disjoint q ranges via read_seed)
    else:
        k = int(workload.split("k")[1][0])
        d_reads = capi.synth_reads_err_device(device, d_text, n_text, nq, L, read_seed, sigma, k, workload.endswith("edit"))
    capi.device_free(device, d_text)
    sym = capi.PinnedArray(nq * L, np.uint8)
    capi.copy_to_host(device, sym.array, d_reads, nq * L)
    capi.device_free(device, d_reads)
    off = capi.PinnedArray(nq + 1, np.uint64)
    off.array[:] = np.arange(nq + 1, dtype=np.uint64) * np.uint64(L)
    log(f"index build {t1 - t0:.1f}s (n={n_text}), reads {time.time() - t1:.1f}s, device image {index.info.device_bytes / 1e9:.2f} GB")
    return index, sym, off


def workload_scheme(workload, L):
    """(scheme, partition, edit) of a k-error workload: optimum(0,k) with a uniform partition (BASELINE configs[2])"""
    from fmb200 import schemes
    if "-k" not in "-" + workload:
        return None, None, False
    k = int(workload.split("k")[1][0])
    sch = schemes.optimum(0, k)
    return sch, schemes.uniform_partition(sch[0].shape[1], L), workload.endswith("edit")


def reference_index(index, sigma=5):
    """host-side reference index over exactly the same BWT bytes and samples (BiFMIndex(bwt, bwtRev, SparseArray))"""
    from oracle.pyoracle import Ref
    t0 = time.time()
    bwt, rev, bm, sq, sp = index.export()
    t1 = time.time()
    ref = Ref.from_bwt(sigma, bwt, rev, bm, sq, sp)
    log(f"reference index: export {t1 - t0:.1f}s, construct {time.time() - t1:.1f}s")
    return ref


def cpu_search_locate(ref, sym, off, b, e, threads, L, scheme=None, partition=None, edit=False, want_locs=False):
    """the reference's search (search_no_errors::search or search_ng26::search<Edit> with the given scheme) + LocateLinear on
    reads [b,e), queries sharded over `threads` std::threads; returns (seconds, n_located[, located rows])"""
    s = sym[b * L: e * L]
    o = (off[b: e + 1] - off[b]).astype(np.uint64)
    hits = ref.search_exact(s, o, threads=threads) if scheme is None else ref.search_ng26(s, o, scheme, partition, edit, threads=threads)
    t = ref.last_seconds
    locs = ref.locate(hits, threads=threads)
    t += ref.last_seconds
    return (t, len(locs), locs) if want_locs else (t, len(locs))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="exact", choices=WORKLOADS,
                    help="exact = BASELINE configs[1] (the headline); the others are configs[2] / configs[3]")
    ap.add_argument("--text", type=float, default=None, help="text length in symbols (default 3 Gbp; 1 Gaa for protein)")
    ap.add_argument("--reads", type=float, default=None, help="reads per GPU (default 10 M; 250 k for locate-heavy)")
    ap.add_argument("--read-len", type=int, default=None, help="default 150 (20 for locate-heavy)")
    ap.add_argument("--rate", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = args.workload
    protein = wl.startswith("protein")
    sigma = 21 if protein else 5
    n_text = int(args.text) if args.text else (1_000_000_000 if protein else 3_000_000_000)
    nq = int(args.reads) if args.reads else (250_000 if wl == "locate-heavy" else (1_000_000 if protein else 10_000_000))
    L = args.read_len if args.read_len else (20 if wl == "locate-heavy" else (50 if protein else 150))
    kk = wl.split("k")[1][0] if "-k" in "-" + wl else "0"
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0

    # stdout carries exactly ONE line, the JSON result: everything libraries print (NCCL's version banner goes to stdout) is
    # sent to stderr by pointing fd 1 at fd 2 for the duration of the run; the result is written to the saved descriptor
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(result_fd, (json.dumps(line) + "\n").encode())

    # host threads of the pipelined one-call path: the library default (6) assumes the host to itself; with one rank per GPU the
    # ranks share the host cores
    os.environ.setdefault("FMB_E2E_THREADS", str(max(2, min(6, (os.cpu_count() or 6) // max(world, 1)))))

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

    scheme, partition, edit = workload_scheme(wl, L)
    what = {"exact": "exact search + locate", "locate-heavy": "exact search + locate of repeat-family 20-mers (many hits per read)"}.get(
        wl, f"k<={kk} {'edit' if edit else 'hamming'} search (optimum scheme, uniform partition) + locate")
    src = "windows of a 1 kbp unit present in 1000 copies" if wl == "locate-heavy" else (
        "reads copied from the text" if wl == "exact" else f"reads copied from the text with 0..{kk} planted {'edits' if edit else 'substitutions'} each")
    unit = "aa" if protein else "bp"
    workload = (f"{what}, {nq} x {L}{unit} {src}, synthetic {n_text} {unit} {'protein' if protein else 'DNA'} BiFMIndex (sigma {sigma}, "
                f"sampling rate {args.rate}), per GPU")
    metric = {"exact": "queries/s (150bp exact search + locate, 3 Gbp index)"}.get(
        wl, f"queries/s ({wl}, search + locate, {'1 Gaa' if protein else '3 Gbp'} index)")
    # the SAME text on every rank (the index is replicated per GPU), different reads per rank (queries are sharded)
    index, sym, off = build_workload(fmb, device, n_text, nq, L, args.rate, seed=3, workload=wl, read_seed=4 + 1000 * rank)
    threads = os.cpu_count() or 1
    ref_name = "search_no_errors::search (batched)" if scheme is None else f"search_ng26::search<{'true' if edit else 'false'}>(optimum(0,{kk}))"

    # ------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        ref = reference_index(index, sigma)
        # size the per-step sample so that W+K steps take about 2 minutes in total
        probe = min(nq, 5000 if scheme is not None or wl == "locate-heavy" else 20000)
        t, _ = cpu_search_locate(ref, sym.array, off.array, 0, probe, threads, L, scheme, partition, edit)
        rate_qs = probe / max(t, 1e-9)
        per_step = int(min(nq, max(probe, rate_qs * 120.0 / (W + K))))
        for _ in range(W):
            cpu_search_locate(ref, sym.array, off.array, 0, per_step, threads, L, scheme, partition, edit)
        tot = 0.0
        for _ in range(K):
            t, nloc = cpu_search_locate(ref, sym.array, off.array, 0, per_step, threads, L, scheme, partition, edit)
            tot += t
        qps = per_step * K / tot
        sample = f"first {per_step} of the {nq} reads per step, {ref_name} + LocateLinear, {threads} threads"
        line = {"impl": "reference", "metric": metric, "value": qps, "unit": "queries/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * tot / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": {"workload": workload},
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "reference", "sample": sample},
                "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------------------------------------------
    stream = torch.cuda.current_stream()
    capi.index_set_stream(index, stream.cuda_stream)
    queries = index.upload(sym.array, off.array)          # resident in HBM before the timed region

    def search():
        return index.search_exact(queries) if scheme is None else index.search_scheme(queries, scheme, partition, edit)

    def step():
        res = search()
        loc = index.locate(res)
        return res, loc

    alg_lookups, one_symbol_ms = None, None
    if scheme is None:
        # algorithmic work of the reference algorithm on this batch (occ-block lookups of one-symbol steps),
        # counted once, untimed, by the one-symbol kernel; the timed steps use the default (two-symbol + jump) kernel
        index.set_exact_mode(1)
        res = index.search_exact(queries)
        alg_lookups = res.stats.occ_lookups
        one_symbol_ms = res.stats.main_kernel_ms
        del res
        index.set_exact_mode(0)

    # clocks / throttle reasons are sampled from the warm-up steps to the end of the e2e loop: the timed region itself lasts only
    # tens of milliseconds, less than one nvidia-smi sampling period
    sampler = ClockSampler(device)
    sampler.start()
    for _ in range(W):
        res, loc = step()
    n_hits, n_locs = len(res), len(loc)
    del res, loc

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = capi.kernel_launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    search_ms, locate_ms, st_s, st_l = [], [], None, None
    for _ in range(K):
        res, loc = step()
        search_ms.append(res.stats.main_kernel_ms)
        locate_ms.append(loc.stats.main_kernel_ms)
        st_s, st_l = res.stats, loc.stats
        del res, loc
    ev1.record(stream)
    barrier()
    launches = capi.kernel_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = nq * world / (ms_per_step * 1e-3)

    # ---- end to end through the C-ABI with host buffers ------------------------------------------------------
    out = capi.PinnedArray(max(n_locs, nq) + 1024, capi.LOC32_DTYPE)
    for _ in range(2):
        index.search_and_locate(sym.array, off.array, scheme=scheme, partition=partition, edit=edit, out=out.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        locs, _st = index.search_and_locate(sym.array, off.array, scheme=scheme, partition=partition, edit=edit, out=out.array)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_qps = nq * world * K / e2e_s
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed steps + e2e loop"
    h2d = nq * L + (nq + 1) * 8
    d2h = len(locs) * 16

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------
    # achieved = ALGORITHMIC bytes (the reference algorithm's occ-block lookups x 32 B) / kernel time.
    # exact_search2_kernel answers those lookups with far fewer physical line fetches (k-mer table, two-symbol lines, LF^16
    # jumps), so `frac` can exceed 1; `physical` is the kernel's own traffic (line requests x 128 B) against the same peak.
    peak, peak_src = load_peaks()
    k_ms = float(np.mean(search_ms))
    l_ms = float(np.mean(locate_ms))
    loc_lookups = n_locs * 2 + st_l.lf_steps * 2     # per row: marker test + sample fetch, per LF step: occ block + marker word
    tables = index.info.tables
    shortcut = bool(tables & 32)                     # FMB_TABLE_LOCROW: the walk is precomputed, two fetches per row
    loc_lines = 2 * n_locs if shortcut else (n_locs + st_l.lf_steps + n_locs)
    locate_info = {"kernel": "locate_shortcut_kernel" if shortcut else ("locate_pair_kernel" if tables & 16 else "locate_kernel"),
                   "kernel_ms": l_ms, "lf_steps_per_row": st_l.lf_steps / max(n_locs, 1),
                   "algorithmic_bytes_per_launch": loc_lookups * 32.0, "achieved": loc_lookups * 32.0 / (l_ms * 1e-3) / 1e9,
                   "frac": loc_lookups * 32.0 / (l_ms * 1e-3) / 1e9 / peak,
                   "physical_lines_per_row": loc_lines / max(n_locs, 1), "physical_lines_per_s": loc_lines / (l_ms * 1e-3)}

    def ncu_traffic(kernel, key):
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                ent = json.load(f)[kernel]
            return ent["traffic_bytes"] if ent["workload"] == key else None
        except Exception:
            return None

    if wl == "locate-heavy":
        roofline = {"bound": "hbm", "kernel": locate_info["kernel"], "achieved": locate_info["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": locate_info["frac"], "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": loc_lookups * 32.0,
                    "kernel_ms": l_ms, "located_rows_per_s": n_locs / (l_ms * 1e-3), "lf_steps_per_row": locate_info["lf_steps_per_row"],
                    "physical_lines_per_row": locate_info["physical_lines_per_row"], "physical_lines_per_s": locate_info["physical_lines_per_s"],
                    "search_kernel_ms": k_ms}
    elif scheme is None:
        alg_bytes = alg_lookups * 32.0
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        phys_bytes = st_s.line_requests * 128.0
        roofline = {"bound": "hbm", "kernel": "exact_search2_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": ncu_traffic("exact_search2_kernel", f"{nq} x {L}bp on {n_text} bp"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "lookups_per_query": alg_lookups / nq, "kernel_ms": k_ms,
                    "physical": {"line_requests_per_query": st_s.line_requests / nq, "bytes_per_launch": phys_bytes,
                                 "gbs": phys_bytes / (k_ms * 1e-3) / 1e9, "frac": phys_bytes / (k_ms * 1e-3) / 1e9 / peak,
                                 "lines_per_s": st_s.line_requests / (k_ms * 1e-3), "measured_random_line_ceiling_per_s": 38.4e9},
                    "one_symbol_kernel_ms": one_symbol_ms, "locate_kernel": locate_info,
                    "note": "frac > 1 is possible: algorithmic bytes are those of the reference's one-symbol algorithm; see physical.frac for the kernel's own traffic"}
    else:
        alg_bytes = st_s.occ_lookups * (64.0 if protein else 32.0)       # 64 B per lookup for sigma = 21
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "scheme_search_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic("scheme_search_kernel:" + wl, f"{nq} x {L}bp on {n_text} bp"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "lookups_per_query": st_s.occ_lookups / nq, "extensions_per_query": st_s.extensions / nq,
                    "kernel_ms": k_ms, "frontier_peak_items_per_warp": st_s.frontier_peak,
                    "physical": {"line_requests_per_query": st_s.line_requests / nq, "lines_per_s": st_s.line_requests / (k_ms * 1e-3),
                                 "measured_random_line_ceiling_per_s": 38.4e9},
                    "locate_kernel": locate_info}

    line = {"metric": metric, "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload, "l2_policy": "inputs larger than L2 (index image %.1f GB, reads %.2f GB)" % (index.info.device_bytes / 1e9, nq * L / 1e9),
                       "index_tables": [name for bit, name in ((1, "pair"), (2, "kmer"), (4, "jump"), (8, "jump_rev"), (16, "locblock"), (32, "locrow"), (64, "bikmer"), (128, "jump4"), (256, "jump32")) if tables & bit],
                       "hits_per_step": n_hits, "located_rows_per_step": n_locs},
            "roofline": roofline,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks}

    # ---- CPU baseline: the reference's own search on this box's host cores (rank 0, N = 1 only) ---------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_index(index, sigma)
            probe = min(nq, 5000 if scheme is not None or wl == "locate-heavy" else 20000)
            t, _ = cpu_search_locate(ref, sym.array, off.array, 0, probe, threads, L, scheme, partition, edit)
            sample_n = int(min(nq, max(probe, probe / max(t, 1e-9) * args.cpu_seconds)))
            t, nloc, exp = cpu_search_locate(ref, sym.array, off.array, 0, sample_n, threads, L, scheme, partition, edit, want_locs=True)
            t1, _ = cpu_search_locate(ref, sym.array, off.array, 0, min(sample_n, probe * 2), 1, L, scheme, partition, edit)
            line["cpu_baseline"] = {"value": sample_n / t, "unit": "queries/s", "cores": threads, "kind": "reference",
                                    "sample": f"first {sample_n} of the {nq} reads, {ref_name} + LocateLinear, {threads} threads "
                                              f"({t:.1f}s); 1 thread: {min(sample_n, probe * 2) / t1:.0f} queries/s"}
            # parity on the sample: located rows identical as sorted multisets
            got = np.sort(locs[locs["qidx"] < sample_n], order=["qidx", "seq", "pos", "e"])
            exp32 = np.zeros(len(exp), dtype=capi.LOC32_DTYPE)
            for f in ("qidx", "seq", "pos", "e"):
                exp32[f] = exp[f]
            exp32 = np.sort(exp32, order=["qidx", "seq", "pos", "e"])
            line["parity"] = {"checked_queries": sample_n, "located_rows": int(len(exp32)), "identical": bool(np.array_equal(got, exp32))}
        except Exception as ex:   # the baseline must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": threads, "kind": "reference", "sample": f"failed: {ex}"}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
