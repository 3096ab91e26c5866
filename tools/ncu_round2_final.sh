#!/bin/bash
# ncu evidence for the final kernels of round 2 (run on the GPU box): launch list of the search path of the default bench (reduced reads)
# and a full-set capture of the text kernel on k = 2 edit (the three launches of the first slab)
set -x
CMD="python bench.py --reads 2e6 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scheme_|exact_search|locate_|pack_queries|unpack_queries|text_class|gather_items|hit_lengths|iota_offsets|gather_probe" -c 2500 --csv --log-file gpurun_out/r02_launches_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1
K2="python bench.py --workload k2-edit --reads 2e6 --steps 1 --warmup 3 --no-cpu-baseline"
$K2 > gpurun_out/ncu_k2e_plain.json 2> gpurun_out/ncu_k2e_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:scheme_text_kernel -s 9 -c 3 -o gpurun_out/r02_prof_text_k2e_final -f $K2 > gpurun_out/ncu_k2e.log 2>&1
ls -la gpurun_out/*final*
