for cfg in "16 96" "8 96" "8 32" "5 24"; do
  set -- $cfg
  echo "== DEPTH $1 BUDGET $2"; FMB_NVCC_EXTRA="-DSIM_DEPTH=$1 -DSIM_BUDGET=$2" python fmindex-collection_b200/build.py --force > /dev/null
  python tools/scheme_bench.py 2>&1 | grep "edit" | cut -c1-60
done
python fmindex-collection_b200/build.py --force > /dev/null
