python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== default (kmer 14 + jump)"; python tools/exact_modes.py 2>&1 | grep -E "mode 2|device image" | tail -2
echo "== no jump"; FMB_NO_JUMP=1 python tools/exact_modes.py 2>&1 | grep "mode 2" | tail -1
echo "== jump, kmer 12"; FMB_KMER_K=12 python tools/exact_modes.py 2>&1 | grep "mode 2" | tail -1
echo "== jump, minb1"; FMB_EXACT2_MINB1=1 python tools/exact_modes.py 2>&1 | grep "mode 2" | tail -1
