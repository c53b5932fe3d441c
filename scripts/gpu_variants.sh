cd "$(dirname "$0")/.."
for k in court court_ultra court_lut; do python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; done
timeout 800 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | head
python tests/gpu_parity_report.py 2>&1 | grep -E "court.*worst" 
