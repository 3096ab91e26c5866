for cfg in "3 160" "3 256" "2 256"; do
  set -- $cfg
  echo "== MINB $1 CAP $2"; FMB_NVCC_EXTRA="-DFMB_SCHEME_MINB=$1" python fmindex-collection_b200/build.py --force > /dev/null
  FMB_SCHEME_CAP=$2 python tools/scheme_bench.py 2>&1 | grep "edit" | cut -c1-60
done
python fmindex-collection_b200/build.py --force > /dev/null
