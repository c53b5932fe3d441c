# ncu evidence for the round: launch list of the bench command + full captures of the 4v / BR / Courtemanche kernels
cd "$(dirname "$0")/.."
B="python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4031 --launch-count 60 --csv --log-file gpurun_out/launches_bench4096.csv $B > gpurun_out/ncu_list.log 2>&1
P="python tests/quick_perf.py 4v 4096 2 --nograph"
$P > gpurun_out/plain_4v.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 32 -c 1 -f -o gpurun_out/prof_4v $P > gpurun_out/ncu_4v_full.log 2>&1
P="python tests/quick_perf.py br 4096 2 --nograph"
$P > gpurun_out/plain_br.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 16 -c 1 -f -o gpurun_out/prof_br $P > gpurun_out/ncu_br_full.log 2>&1
P="python tests/quick_perf.py br_exact 4096 2 --nograph"
$P > gpurun_out/plain_brx.log 2>&1 && ncu --set full --clock-control none -k regex:step_kernel -s 16 -c 1 -f -o gpurun_out/prof_br_exact $P > gpurun_out/ncu_brx_full.log 2>&1
P="python tests/quick_perf.py court_ultra 4096 2 --nograph"
$P > gpurun_out/plain_court.log 2>&1 && ncu --set full --clock-control none -k regex:step_kernel -s 4 -c 1 -f -o gpurun_out/prof_court $P > gpurun_out/ncu_court_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
