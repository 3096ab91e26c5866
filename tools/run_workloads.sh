# full-size runs of the non-headline workloads (BASELINE configs[2..4]); one JSON line each under gpurun_out/
for wl in "$@"; do
  python bench.py --workload $wl --steps 3 --warmup 3 --cpu-seconds 6 > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err || { tail -5 gpurun_out/wl_$wl.err; }
  python - <<PY
import json
d=json.load(open("gpurun_out/wl_$wl.json")); r=d["roofline"]
print("$wl", "value %.1f Mq/s" % (d["value"]/1e6), "e2e %.1f" % (d["e2e"]["value"]/1e6), "ms/step %.2f" % d["ms_per_step"], "kernel_ms %.2f" % r["kernel_ms"], "frac %.3f" % r["frac"], "hits", d["config"]["hits_per_step"], "cpu", d.get("cpu_baseline",{}).get("value"), d.get("parity"))
PY
done
