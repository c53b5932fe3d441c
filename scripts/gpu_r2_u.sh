cd "$(dirname "$0")/.."
for m in br_skip court_lut br; do
  timeout 300 python tests/quick_perf.py $m 2048 60 2>&1 | tail -1
  timeout 300 python tests/quick_perf.py $m 2048 60 --nograph 2>&1 | tail -1
  FIB_PDL=0 timeout 300 python tests/quick_perf.py $m 2048 60 --nograph 2>&1 | tail -1
done
