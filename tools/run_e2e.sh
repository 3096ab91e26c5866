for cfg in "19 8" "19 12" "18 12" "18 16"; do set -- $cfg; echo "== chunk 2^$1 threads $2"; FMB_E2E_CHUNK_LOG2=$1 FMB_E2E_THREADS=$2 python tools/e2e_scheme.py k1-hamming 2>&1 | tail -1;  FMB_E2E_CHUNK_LOG2=$1 FMB_E2E_THREADS=$2 python tools/e2e_trace.py 2>&1 | tail -1; done
FMB_E2E_CHUNK_LOG2=19 FMB_E2E_THREADS=8 python tools/e2e_scheme.py k2-edit 2>&1 | tail -1
FMB_E2E_CHUNK_LOG2=20 FMB_E2E_THREADS=3 python tools/e2e_scheme.py k2-edit 2>&1 | tail -1
