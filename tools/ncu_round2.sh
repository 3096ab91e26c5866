#!/bin/bash
# round-2 ncu evidence (run on the GPU box): launch list of the search path of the default bench (reduced reads), full-set captures
# of the text kernel (k = 2 edit, the three launches of one slab), the frontier kernel (k = 1 edit) and the exact kernel
set -x
CMD="python bench.py --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scheme_|exact_search|locate_|pack_queries|unpack_queries|text_class|gather_items|hit_lengths|iota_offsets|gather_probe" -c 2500 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
K2="python bench.py --workload k2-edit --reads 2e6 --steps 1 --warmup 3 --no-cpu-baseline"
$K2 > gpurun_out/ncu_k2e_plain.json 2> gpurun_out/ncu_k2e_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:scheme_text_kernel -s 9 -c 3 -o gpurun_out/r02_prof_text_k2e -f $K2 > gpurun_out/ncu_k2e.log 2>&1
K1="python bench.py --workload k1-edit --reads 2e6 --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:scheme_search_kernel -s 6 -c 1 -o gpurun_out/r02_prof_frontier_k1e -f $K1 > gpurun_out/ncu_k1e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scheme_text_kernel -s 3 -c 1 -o gpurun_out/r02_prof_text_k1e -f $K1 > gpurun_out/ncu_k1e_text.log 2>&1
EX="python bench.py --workload exact --reads 1e7 --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:exact_search2_kernel -s 3 -c 1 -o gpurun_out/r02_prof_exact2 -f $EX > gpurun_out/ncu_exact.log 2>&1
ls -la gpurun_out/*.ncu-rep
