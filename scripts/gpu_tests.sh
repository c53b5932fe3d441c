cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
python __graft_entry__.py --smoke 2>&1 | tail -5
B="python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4011 --launch-count 60 --csv --log-file gpurun_out/launches_bench4096.csv $B > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
