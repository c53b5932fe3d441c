set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python __graft_entry__.py --smoke 2>&1 | tail -8
python tests/gpu_parity_report.py 2>&1 | tail -70
for k in 4v br br_exact br_skip court court_ultra court_lut; do python tests/quick_perf.py $k 4096 10; done
python tests/quick_perf.py 4v 4096 10 --phase
python tests/quick_perf.py 4v 512 200
python tests/quick_perf.py 4v 512 200 --nograph
python tests/quick_perf.py br 512 200
