#!/bin/bash
# Build libscipnp.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/csrc"
OUT="$HERE/scipnp/libscipnp.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p "$HERE/build"
OBJS=""
for f in api ops tv_exact gap_tv_fused solver; do
  o="$HERE/build/$f.o"
  if [ ! -f "$o" ] || [ "$SRC/$f.cu" -nt "$o" ] || \
     [ -n "$(find "$SRC" "$HERE/../include" \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ]; then
    echo "nvcc $f.cu"
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS} -c "$SRC/$f.cu" -o "$o"
  fi
  OBJS="$OBJS $o"
done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a $OBJS -o "$OUT" -cudart static
echo "built $OUT"
