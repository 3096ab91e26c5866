for ff in 1 8 16 24 33; do echo "== ff_min $ff"; FMB_SCHEME_FFMIN=$ff python tools/scheme_bench.py 2>&1 | grep "kernel" ; done
