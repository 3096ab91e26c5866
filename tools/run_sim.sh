for sa in 0 1; do for ff in 16 24 33; do echo "== SIMALL $sa FFMIN $ff"; FMB_SCHEME_SIMALL=$sa FMB_SCHEME_FFMIN=$ff python tools/scheme_bench.py 2>&1 | grep "edit" | cut -c1-60; done; done
