for mb in 2 4; do
  echo "== MINB $mb"; FMB_NVCC_EXTRA="-DFMB_SCHEME_MINB=$mb" python fmindex-collection_b200/build.py --force > /dev/null
  python tools/scheme_bench.py 2>&1 | grep "kernel" | cut -c1-60
done
python fmindex-collection_b200/build.py --force > /dev/null
