# round 2, call A: parity reports (default / wide / strict), tuning variants, suite, full GPU tests, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tests/gpu_parity_report.py --out gpurun_out/r2_parity_report.txt > /dev/null 2>gpurun_out/rep.err; echo "report rc=$?"; tail -3 gpurun_out/rep.err
FIB_SMALL_CELLS=0 python tests/gpu_parity_report.py --out gpurun_out/r2_parity_report_wide.txt > /dev/null 2>&1
python tests/gpu_parity_report.py --strict --out gpurun_out/r2_parity_report_strict.txt > /dev/null 2>&1
FIB_SMALL_CELLS=0 python tests/gpu_parity_report.py --strict --out gpurun_out/r2_parity_report_strict_wide.txt > /dev/null 2>&1
FIB_SMALL_CELLS=0 FIB_B200_LIB=$PWD/build/variants/lib_all2_4.so python tests/gpu_parity_report.py --out gpurun_out/r2_parity_report_wide_all2.txt > /dev/null 2>&1
grep -h "ABOVE\|FAIL\|^#" gpurun_out/r2_parity_report*.txt | head -120
echo "== variants (4096^2, Gcell-steps/s)"
for k in br br_exact br_skip court court_ultra court_lut; do python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; done
for v in all2_4 all2_3; do echo "-- $v"; FIB_B200_LIB=$PWD/build/variants/lib_$v.so python tests/quick_perf.py court_ultra 4096 6 2>&1 | tail -1; done
for v in br5 br6 br8; do echo "-- $v"; for k in br br_exact; do FIB_B200_LIB=$PWD/build/variants/lib_$v.so python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; done; done
for v in lut3 lut5; do echo "-- $v"; FIB_B200_LIB=$PWD/build/variants/lib_$v.so python tests/quick_perf.py court_lut 4096 6 2>&1 | tail -1; done
python scripts/suite.py r2a 2>&1 | tail -14
timeout 2400 python -m pytest tests -m gpu -q -rf --timeout 1200 > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2a_tests.log
python __graft_entry__.py --smoke 2>&1 | tail -6
