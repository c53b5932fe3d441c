# what the driver runs at round end, in one go, plus the evidence files regenerated from the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/spiral_report.txt
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
FIB_SPIRAL_REPORT=gpurun_out/spiral_report.txt timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tests/gpu_parity_report.py --out gpurun_out/r2_parity_report.txt > /dev/null 2>&1; echo "report rc=$?"
FIB_SMALL_CELLS=0 FIB_PERSIST=0 python tests/gpu_parity_report.py --out gpurun_out/r2_parity_report_wide.txt > /dev/null 2>&1; echo "wide rc=$?"
python tests/gpu_parity_report.py --strict --out gpurun_out/r2_parity_report_strict.txt > /dev/null 2>&1; echo "strict rc=$?"
FIB_SMALL_CELLS=0 FIB_PERSIST=0 python tests/gpu_parity_report.py --strict --out gpurun_out/r2_parity_report_strict_wide.txt > /dev/null 2>&1; echo "strict wide rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; python -c "import json;d=json.load(open('gpurun_out/final_ref.json'));print('reference arm', d['value'], d['cpu_baseline']['cores'], d['steps'])"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ours.json 2> gpurun_out/final_ours.err; echo "stdout lines: $(wc -l < gpurun_out/final_ours.json)"; python -c "
import json;d=json.load(open('gpurun_out/final_ours.json'))
need=['metric','value','unit','n_gpus','steps','warmup','ms_per_step','higher_is_better','scaling','vs_baseline','dtype','data','config','e2e','gpu_launches','clocks','roofline','cpu_baseline']
print('missing keys:', [k for k in need if k not in d])
print('ours', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
for e in d['suite']: print('  %-90s %8.1f  frac %.3f' % (e['case'][:90], e['value'], e['frac']))"
