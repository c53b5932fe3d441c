# what the driver runs at round end, in one go
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; python -c "import json;d=json.load(open('gpurun_out/final_ref.json'));print('reference arm', d['value'], d['cpu_baseline']['cores'], d['steps'])"
python bench.py > gpurun_out/final_ours.json 2> gpurun_out/final_ours.err; python -c "
import json;d=json.load(open('gpurun_out/final_ours.json'))
need=['metric','value','unit','n_gpus','steps','warmup','ms_per_step','higher_is_better','scaling','vs_baseline','dtype','data','config','e2e','gpu_launches','clocks','roofline','cpu_baseline']
print('missing keys:', [k for k in need if k not in d])
print('ours', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
