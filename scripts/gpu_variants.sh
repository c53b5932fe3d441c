cd "$(dirname "$0")/.."
for pdl in 1 0 1 0; do
  export FIB_PDL=$pdl; echo "=== FIB_PDL=$pdl"
  for k in 4v br court_ultra; do python tests/quick_perf.py $k 512 300 | tail -1; done
  python tests/quick_perf.py 4v 4096 6 | tail -1; python tests/quick_perf.py br 4096 6 | tail -1
  python tests/quick_perf.py 4v 512 300 --nograph | tail -1
done
unset FIB_PDL
timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -3
