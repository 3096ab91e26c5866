#!/bin/bash
# throughput of the five headline workloads against the HBM budget of the index image (run on the GPU box)
for gb in 16 40 64 78 102 0; do
  python bench.py --image-gb $gb --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_budget_${gb}.json 2> gpurun_out/r02_budget_${gb}.log || echo "budget $gb failed"
  grep "^\[bench\]" gpurun_out/r02_budget_${gb}.log | grep -v reference
done
