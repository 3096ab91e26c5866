ncu --set full --clock-control none --import-source on -k regex:scheme_search_kernel -s 3 -c 1 -o gpurun_out/prof_scheme_k2e_v2 -f \
    python bench.py --workload k2-edit --reads 1e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k2e_v2.log 2>&1
ncu -i gpurun_out/prof_scheme_k2e_v2.ncu-rep --page source --csv --print-source cuda > gpurun_out/src_k2e.csv 2>/dev/null || ncu -i gpurun_out/prof_scheme_k2e_v2.ncu-rep --page source --csv > gpurun_out/src_k2e.csv
ls -la gpurun_out/src_k2e.csv
