# ncu evidence for one round (run on the GPU box AFTER the plain bench exited 0): launch list + full-set captures
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:exact_search2_kernel -s 3 -c 1 -o gpurun_out/prof_exact2 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_exact2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:locate_shortcut_kernel -s 3 -c 1 -o gpurun_out/prof_locate3 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_locate3.log 2>&1
python bench.py --workload k1-hamming --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1h_plain.json 2> gpurun_out/ncu_k1h_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:scheme_search_kernel -s 3 -c 1 -o gpurun_out/prof_scheme_k1h -f \
    python bench.py --workload k1-hamming --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1h.log 2>&1
python bench.py --workload k2-edit --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k2e_plain.json 2> gpurun_out/ncu_k2e_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:scheme_search_kernel -s 3 -c 1 -o gpurun_out/prof_scheme_k2e -f \
    python bench.py --workload k2-edit --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k2e.log 2>&1
ls -la gpurun_out/ | tail -15
