# builds tuning variants of libfibb200.so into build/variants/ (they travel to the GPU box with the
# snapshot; select one with FIB_B200_LIB=...)   usage: scripts/build_variants.sh name "-DFOO=1 -DBAR=2" ...
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared $flags \
      -o build/variants/lib_$name.so fib_tf_b200/csrc/fib_capi.cu -ldl 2>&1 | grep -E "error" ; echo "built $name" ) &
done
wait
