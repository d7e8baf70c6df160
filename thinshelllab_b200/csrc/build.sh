#!/bin/bash
# Builds libtsl.so in-tree for sm_100a (the .so travels to the GPU box with the gpurun snapshot).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libtsl.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
mkdir -p "$HERE/_obj"
pids=()
for f in tsl_physics tsl_contact tsl_linalg tsl_mg tsl_api tsl_dist tsl_dense tsl_assembly; do
  if [ ! -f "$HERE/_obj/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/_obj/$f.o" ] || [ -n "$(find "$HERE" "$HERE/../../include" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$HERE/_obj/$f.o" 2>/dev/null)" ]; then
    EXTRA=""; [ "$f" = tsl_contact ] && EXTRA="--fmad=false"   # contact decisions must match the fp64 oracle bit for bit
    $NVCC $FLAGS $EXTRA ${TSL_PTXAS_V:+-Xptxas -v} -c "$HERE/$f.cu" -o "$HERE/_obj/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT" "$HERE"/_obj/tsl_physics.o "$HERE"/_obj/tsl_contact.o "$HERE"/_obj/tsl_linalg.o "$HERE"/_obj/tsl_mg.o "$HERE"/_obj/tsl_api.o "$HERE"/_obj/tsl_dist.o "$HERE"/_obj/tsl_dense.o "$HERE"/_obj/tsl_assembly.o -ldl -gencode arch=compute_100a,code=sm_100a
echo "built $OUT"
