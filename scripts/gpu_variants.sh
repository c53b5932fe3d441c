cd "$(dirname "$0")/.."
for k in 4v br br_exact br_skip court court_ultra court_lut; do python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; done
python tests/quick_perf.py 4v 4096 6 --phase | tail -1
python tests/quick_perf.py br 4096 6 --phase | tail -1
for r in 1 2; do export FIB_FORCE_R=$r; echo "R=$r"; python tests/quick_perf.py court 4096 6 | tail -1; python tests/quick_perf.py court_lut 4096 6 | tail -1; done; unset FIB_FORCE_R
for k in 4v br br_skip court_ultra; do python tests/quick_perf.py $k 512 200 | tail -1;  python tests/quick_perf.py $k 2048 20 | tail -1; done
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
python tests/gpu_parity_report.py 2>&1 | grep -E "worst|WORST"
