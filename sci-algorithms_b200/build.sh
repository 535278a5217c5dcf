#!/bin/bash
# Build libscipnp.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
# Units are compiled in parallel; an object is rebuilt when its source or any header is newer.
# SCIPNP_FAST=1 builds the fused kernel for C = 8 and C = 24 only (development shortcut).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/csrc"
OUT="$HERE/scipnp/libscipnp.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
if [ -n "$SCIPNP_FAST" ]; then FLAGS="$FLAGS -DSCIPNP_FUSED_FAST_BUILD"; fi
mkdir -p "$HERE/build"
UNITS="api ops tv_exact tv_matlab gap_tv_fused fused_inst_r2 fused_inst_r3 fused_inst_r4 fused_inst_r4c gap_tv_ws ws_inst_r2 ws_inst_r3 ws_inst_r4 solver"
OBJS=""
PIDS=""
STAMP="$HERE/build/.flags"
if [ ! -f "$STAMP" ] || [ "$(cat "$STAMP")" != "$FLAGS" ]; then rm -f "$HERE"/build/*.o; echo "$FLAGS" > "$STAMP"; fi
for f in $UNITS; do
  o="$HERE/build/$f.o"
  if [ ! -f "$o" ] || [ "$SRC/$f.cu" -nt "$o" ] || \
     [ -n "$(find "$SRC" "$HERE/../include" \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ]; then
    echo "nvcc $f.cu"
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS} -c "$SRC/$f.cu" -o "$o" &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $o"
done
for p in $PIDS; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a $OBJS -o "$OUT" -cudart static
echo "built $OUT"
