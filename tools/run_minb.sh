python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "4 160" "3 256" "3 160"; do
  set -- $cfg
  echo "== MINB $1 CAP $2"; FMB_NVCC_EXTRA="-DFMB_SCHEME_MINB=$1" python fmindex-collection_b200/build.py --force > /dev/null
  FMB_SCHEME_CAP=$2 python tools/scheme_bench.py 2>&1 | grep "kernel" | cut -c1-60
done
python fmindex-collection_b200/build.py --force > /dev/null
python tools/dbg_counters.py 2>&1 | tail -2
