# ncu evidence for the round: launch list of the bench command + full captures of the 4v (one and two
# steps per launch) / BR / Courtemanche kernels.  Two parts (a gpurun call brings back <= 64 MiB):
#   bash scripts/gpu_ncu.sh 1   -> launch list, 4v, 4v fused, BR polynomial gates
#   bash scripts/gpu_ncu.sh 2   -> BR exact gates, Courtemanche all-state and LUT, the persistent kernel (4v, BR)
cd "$(dirname "$0")/.."
PART=${1:-1}
full() {   # name, kernel regex, launches to skip, command...
  name=$1; shift; rx=$1; shift; skip=$1; shift
  "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_${name}_full.log 2>&1
  # a call brings back <= 64 MiB: export the pages that get read and drop reports that are too big to travel
  rep=gpurun_out/prof_$name.ncu-rep
  if [ -f $rep ]; then
    ncu -i $rep --page raw --csv > gpurun_out/prof_$name.raw.csv 2>/dev/null
    ncu -i $rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/prof_$name.source.csv.gz
    ncu -i $rep --page details 2>/dev/null | gzip -9 > gpurun_out/prof_$name.details.txt.gz
    [ $(stat -c %s $rep) -gt 12000000 ] && rm -f $rep
  fi
}
if [ "$PART" = 0 ] || [ "$PART" = 1 ]; then
  B="python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu --no-suite"
  $B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --launch-count 900 --csv --log-file gpurun_out/launches_bench4096.csv $B > gpurun_out/ncu_list.log 2>&1
  [ "$PART" = 0 ] && { tail -c 300 gpurun_out/plain_bench.log; wc -l gpurun_out/launches_bench4096.csv; exit 0; }
  FIB_STEPS_PER_LAUNCH=1 full 4v step_kernel 32 python tests/quick_perf.py 4v 4096 2 --nograph
  full 4v_fused fused2 10 python tests/quick_perf.py 4v 4096 2 --nograph
  full br step_kernel 16 python tests/quick_perf.py br 4096 2 --nograph
else
  full br_exact step_kernel 16 python tests/quick_perf.py br_exact 4096 2 --nograph
  full court step_kernel 4 python tests/quick_perf.py court_ultra 4096 2 --nograph
  full court_lut step_kernel 4 python tests/quick_perf.py court_lut 4096 2 --nograph
  full persist_4v persist_kernel 1 timeout 300 python scripts/persist_probe.py 4v 64
  full persist_br persist_kernel 1 timeout 300 python scripts/persist_probe.py br 64
fi
ls -la gpurun_out/ | grep prof_; du -sh gpurun_out
