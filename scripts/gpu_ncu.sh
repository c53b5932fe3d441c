# ncu evidence for the round: launch list of the bench command + full captures of the 4v (one and two
# steps per launch) / BR / Courtemanche kernels.  Two parts (a gpurun call brings back <= 64 MiB):
#   bash scripts/gpu_ncu.sh 1   -> launch list, 4v, 4v fused, BR polynomial gates
#   bash scripts/gpu_ncu.sh 2   -> BR exact gates, Courtemanche all-state
cd "$(dirname "$0")/.."
PART=${1:-1}
full() {   # name, kernel regex, launches to skip, command...
  name=$1; shift; rx=$1; shift; skip=$1; shift
  "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_${name}_full.log 2>&1
}
if [ "$PART" = 1 ]; then
  B="python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu"
  $B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4011 --launch-count 60 --csv --log-file gpurun_out/launches_bench4096.csv $B > gpurun_out/ncu_list.log 2>&1
  FIB_STEPS_PER_LAUNCH=1 full 4v step_kernel 32 python tests/quick_perf.py 4v 4096 2 --nograph
  full 4v_fused fused2 10 python tests/quick_perf.py 4v 4096 2 --nograph
  full br step_kernel 16 python tests/quick_perf.py br 4096 2 --nograph
else
  full br_exact step_kernel 16 python tests/quick_perf.py br_exact 4096 2 --nograph
  full court step_kernel 4 python tests/quick_perf.py court_ultra 4096 2 --nograph
fi
ls -la gpurun_out/*.ncu-rep
