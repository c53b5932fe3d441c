# round 2, call B: persistent kernel (guarded by timeouts), tuning sweep, tests, reports
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== persistent kernel tests"
timeout 600 python -m pytest tests/test_gpu_persist.py -q -x --timeout 300 2>&1 | tail -25
echo "== 512^2 configs: persistent vs one launch per step"
python scripts/suite.py r2b 2 2>&1 | tail -3
FIB_PERSIST=0 python scripts/suite.py r2b_nopersist 2 2>&1 | tail -3
echo "== tuning sweep (4096^2, Gcell-steps/s)"
for v in default v1 v2 v3 v4; do echo "-- $v"; for k in br br_exact br_skip court court_ultra court_lut; do
  if [ $v = default ]; then python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; else FIB_B200_LIB=$PWD/build/variants/lib_$v.so python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; fi; done; done
echo "== full GPU tests"
timeout 2400 python -m pytest tests -m gpu -q -rf --timeout 1200 > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?"
grep -E "^E  .*(Error|assert)|^FAILED|passed|failed" gpurun_out/r2b_tests.log | head -40
python tests/gpu_parity_report.py --out gpurun_out/r2_parity_report.txt > /dev/null 2>&1
FIB_SMALL_CELLS=0 python tests/gpu_parity_report.py --out gpurun_out/r2_parity_report_wide.txt > /dev/null 2>&1
python tests/gpu_parity_report.py --strict --out gpurun_out/r2_parity_report_strict.txt > /dev/null 2>&1
grep -h "FAIL\|^#" gpurun_out/r2_parity_report.txt gpurun_out/r2_parity_report_wide.txt gpurun_out/r2_parity_report_strict.txt | head -30
python __graft_entry__.py --smoke 2>&1 | tail -6
