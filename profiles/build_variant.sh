#!/bin/bash
# build an experimental variant of the engine library: profiles/build_variant.sh <name> [-DFLAG=V ...]
# -> magent_b200/lib/variants/libmagent_<name>.so ; select it with MAGENT_B200_LIB=<path>
set -e
root="$(cd "$(dirname "$0")/.." && pwd)"
name="$1"; shift
out="$root/magent_b200/lib/variants"; mkdir -p "$out"
src="$root/magent_b200/csrc"
for f in engine.cc shim.cc host_expand.cc; do
  /usr/bin/g++ -O3 -std=c++17 -fPIC -fvisibility=hidden -pthread "$@" -c "$src/$f" -o "$out/${name}_${f%.*}.o"
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr "$@" -x cu -c "$src/backend_cuda.cu" -o "$out/${name}_backend_cuda.o"
/usr/local/cuda/bin/nvcc -shared -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -Xlinker -Bsymbolic -Xcompiler -pthread \
  -o "$out/libmagent_$name.so" "$out/${name}_engine.o" "$out/${name}_shim.o" "$out/${name}_host_expand.o" "$out/${name}_backend_cuda.o"
rm -f "$out/${name}_"*.o
echo "$out/libmagent_$name.so"
