# ncu capture of one kernel family: bash scripts/gpu_ncu.sh <quick_perf model> <skip> <tag> [lib]
cd "$(dirname "$0")/.."
M=$1; S=$2; TAG=$3
[ -n "$4" ] && export FIB_B200_LIB=$PWD/$4
P="python tests/quick_perf.py $M 4096 2 --nograph"
$P > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $S -c 1 -f -o gpurun_out/prof_$TAG $P > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out/
