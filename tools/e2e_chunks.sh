#!/bin/bash
# end-to-end rate of the one-call path against the chunk size (FMB_E2E_CHUNK_LOG2), byte and packed input, 10 M x 150 bp on 3 Gbp
for wl in k1-hamming k2-edit; do for lg in 20 21 22; do echo "== $wl chunk 2^$lg"; FMB_E2E_CHUNK_LOG2=$lg python tools/e2e_trace.py $wl 1e7 3e9 2>&1 | grep ": call " | sed -n '4p;8p'; done; done
