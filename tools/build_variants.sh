#!/bin/bash
# Kernel experiments: builds libb200ida variants with different -D switches into variants/ (git-ignored; they travel to
# the GPU box with the snapshot).  usage: tools/build_variants.sh name1:"-DFLAG=1 -DOTHER=2" name2:"..." ...
# Each variant is loaded through the B200IDA_LIB override of rust_ida_b200/_lib.py (tools/bench_variants.py).
set -e
cd "$(dirname "$0")/../rust_ida_b200/csrc"
mkdir -p ../../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVFLAGS="$ARCH -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
[ -f linalg_batched.o ] || nvcc $NVFLAGS -c linalg_batched.cu -o linalg_batched.o
[ -f newton_batched.o ] || nvcc $NVFLAGS -c newton_batched.cu -o newton_batched.o
pids=()
for spec in "$@"; do
    name="${spec%%:*}"
    flags="${spec#*:}"
    (
        nvcc $NVFLAGS $flags -c ida_engine.cu -o ../../variants/ida_engine_$name.o
        nvcc $ARCH -shared -o ../../variants/libb200ida_$name.so ../../variants/ida_engine_$name.o linalg_batched.o newton_batched.o --cudart static
        rm -f ../../variants/ida_engine_$name.o
        echo "built variants/libb200ida_$name.so ($flags)"
    ) &
    pids+=($!)
    # at most 4 compilations at a time
    if [ ${#pids[@]} -ge 4 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
done
wait
