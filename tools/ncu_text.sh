#!/bin/bash
# full-set ncu capture of the text kernel on the k = 2 edit workload (2 M reads), all launches of one search call
K2="python bench.py --workload k2-edit --reads 2e6 --steps 1 --warmup 3 --no-cpu-baseline"
$K2 > gpurun_out/ncu_k2e_plain.json 2> gpurun_out/ncu_k2e_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:scheme_text_kernel -s 9 -c 3 -o gpurun_out/r02_prof_text_k2e_final -f $K2 > gpurun_out/ncu_k2e.log 2>&1
ls -la gpurun_out/r02_prof_text_k2e_final.ncu-rep
