set -x
cd "$(dirname "$0")/.."
python tests/gpu_parity_report.py 2>&1 | grep -E "worst|WORST" 
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40
