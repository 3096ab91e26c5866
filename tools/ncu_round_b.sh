# second half of ncu_round.sh: full-set capture of the scheme kernel (k = 1 Hamming, 2 M reads)
set -x
python bench.py --workload k1-hamming --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1h_plain.json 2> gpurun_out/ncu_k1h_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:scheme_search_kernel -s 3 -c 1 -o gpurun_out/prof_scheme_k1h -f \
    python bench.py --workload k1-hamming --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1h.log 2>&1
