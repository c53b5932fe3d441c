cd "$(dirname "$0")/.."
for k in 4v br court_ultra; do python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; python tests/quick_perf.py $k 4096 6 --phase 2>&1 | tail -1; done
python tests/quick_perf.py 4v 512 200 --phase | tail -1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
