for wl in k1-hamming k1-edit k2-hamming k2-edit; do
  python bench.py --workload $wl --reads 1e6 --steps 3 --warmup 3 --cpu-seconds 5 > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err || { tail -5 gpurun_out/wl_$wl.err; }
  python - <<PY
import json
d=json.load(open("gpurun_out/wl_$wl.json")); r=d["roofline"]
print("$wl", "value %.1f Mq/s" % (d["value"]/1e6), "e2e %.1f" % (d["e2e"]["value"]/1e6), "kernel_ms %.2f" % r["kernel_ms"], "locate_ms %.2f" % r["locate_kernel"]["kernel_ms"], "frac %.3f" % r["frac"], "hits", d["config"]["hits_per_step"], "cpu", d.get("cpu_baseline",{}).get("value"), d.get("parity"))
PY
done
python bench.py --workload locate-heavy --steps 3 --warmup 3 --cpu-seconds 5 > gpurun_out/wl_locate.json 2> gpurun_out/wl_locate.err || tail -5 gpurun_out/wl_locate.err
cat gpurun_out/wl_locate.json
