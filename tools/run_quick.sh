python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/scheme_bench.py 2>&1 | grep "kernel" | cut -c1-60
python tools/dbg_counters.py 2>&1 | tail -2
