cd "$(dirname "$0")/.."
for k in 4v br br_exact br_skip court court_ultra court_lut; do python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; done
timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -2
