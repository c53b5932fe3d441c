cd "$(dirname "$0")/.."
for v in a b c d e; do
  export FIB_B200_LIB=$PWD/build/variants/lib_$v.so
  echo "=== $v"
  for k in br_exact court court_ultra court_lut; do python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; done
done
