cd "$(dirname "$0")/.."
mkdir -p gpurun_out
name=persist_4v_b
timeout 300 python scripts/persist_probe.py 4v 64 > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:persist_kernel -s 1 -c 1 -f -o gpurun_out/prof_$name timeout 300 python scripts/persist_probe.py 4v 64 > gpurun_out/ncu_${name}_full.log 2>&1
rep=gpurun_out/prof_$name.ncu-rep
ncu -i $rep --page raw --csv > gpurun_out/prof_$name.raw.csv 2>/dev/null
ncu -i $rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/prof_$name.source.csv.gz
rm -f $rep
ls -la gpurun_out | grep $name
