set -x
python bench.py --workload k1-hamming --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1h_plain.json 2> gpurun_out/ncu_k1h_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:scheme_search_kernel -s 3 -c 1 -o gpurun_out/prof_scheme_k1h -f \
    python bench.py --workload k1-hamming --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1h.log 2>&1
python bench.py --workload k2-edit --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k2e_plain.json 2> gpurun_out/ncu_k2e_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:scheme_search_kernel -s 3 -c 1 -o gpurun_out/prof_scheme_k2e -f \
    python bench.py --workload k2-edit --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k2e.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
