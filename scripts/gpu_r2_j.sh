# round 2, call J: batched persistent launches, full tests, parity reports, suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/persist_probe.py 4v 300; python scripts/persist_probe.py br 300
echo "== full GPU tests"
timeout 2400 python -m pytest tests -m gpu -q -rf --timeout 1200 > gpurun_out/r2j_tests.log 2>&1; echo "pytest rc=$?"
grep -E "^E  .*(Error|assert)|^FAILED|passed|failed" gpurun_out/r2j_tests.log | cut -c1-220 | head -30
python tests/gpu_parity_report.py --out gpurun_out/r2_parity_report.txt > /dev/null 2>&1
FIB_SMALL_CELLS=0 FIB_PERSIST=0 python tests/gpu_parity_report.py --out gpurun_out/r2_parity_report_wide.txt > /dev/null 2>&1
python tests/gpu_parity_report.py --strict --out gpurun_out/r2_parity_report_strict.txt > /dev/null 2>&1
FIB_SMALL_CELLS=0 FIB_PERSIST=0 python tests/gpu_parity_report.py --strict --out gpurun_out/r2_parity_report_strict_wide.txt > /dev/null 2>&1
grep -h "FAIL\|^#" gpurun_out/r2_parity_report*.txt | head -30
python __graft_entry__.py --smoke 2>&1 | tail -6
python bench.py --size 8192 --steps 5 --warmup 3 > gpurun_out/r2j_bench8192.json 2> gpurun_out/r2j_bench8192.err; echo "bench rc=$?"; tail -2 gpurun_out/r2j_bench8192.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench8192.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'dram_frac',d['roofline'].get('dram_frac'))
for s in d.get('suite',[]): print('  %-90s %8.2f  %6.2f us/step frac %.3f'%(s['case'][:90],s.get('value',0),s.get('us_per_time_step',0),s.get('frac',0)))
PY
