#!/bin/bash
# the other BASELINE configs, one JSON line each under gpurun_out/ (run on the GPU box)
set -x
S="--steps 10 --warmup 3"
python bench.py --workload c1 $S > gpurun_out/r02_bench_c1_N1.json 2> gpurun_out/r02_bench_c1_N1.log
python bench.py --workload locate-heavy --rate 16 $S > gpurun_out/r02_bench_locate-heavy_rate16_N1.json 2> gpurun_out/r02_bench_locate-heavy_rate16_N1.log
python bench.py --workload locate-heavy --rate 32 $S > gpurun_out/r02_bench_locate-heavy_rate32_N1.json 2> gpurun_out/r02_bench_locate-heavy_rate32_N1.log
python bench.py --workload protein-k1-hamming $S > gpurun_out/r02_bench_protein-k1-hamming_N1.json 2> gpurun_out/r02_bench_protein-k1-hamming_N1.log
python bench.py --workload protein-k1-edit $S > gpurun_out/r02_bench_protein-k1-edit_N1.json 2> gpurun_out/r02_bench_protein-k1-edit_N1.log
python bench.py --workload repeat150 $S > gpurun_out/r02_bench_repeat150_N1.json 2> gpurun_out/r02_bench_repeat150_N1.log
grep -h "^\[bench\]" gpurun_out/r02_bench_c1_N1.log gpurun_out/r02_bench_locate-heavy_rate16_N1.log gpurun_out/r02_bench_locate-heavy_rate32_N1.log gpurun_out/r02_bench_protein-k1-hamming_N1.log gpurun_out/r02_bench_protein-k1-edit_N1.log gpurun_out/r02_bench_repeat150_N1.log | grep -v "reference index"
