for cfg in "18 3" "20 3" "20 4" "21 3" "19 6"; do set -- $cfg; echo "== chunk 2^$1 threads $2"; FMB_E2E_CHUNK_LOG2=$1 FMB_E2E_THREADS=$2 python tools/e2e_scheme.py k1-hamming 2>&1 | tail -1; done
