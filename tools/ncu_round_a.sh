# first half of ncu_round.sh: launch list + full-set captures of the exact-search and locate kernels
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:exact_search2_kernel -s 3 -c 1 -o gpurun_out/prof_exact2 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_exact2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:locate_shortcut_kernel -s 3 -c 1 -o gpurun_out/prof_locate3 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_locate3.log 2>&1
ls -la gpurun_out/ | tail -8
