#!/bin/bash
# Build an A/B variant of libmst_b200.so with extra -D flags into ml_music_style_transfer_b200/variants/<name>/ .
# usage: tools/build_variant.sh <name> [-DFOO=1 ...];  run with LD_PRELOAD=<that .so> python bench.py ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/ml_music_style_transfer_b200/variants/$name
mkdir -p "$out"
cd "$root/ml_music_style_transfer_b200/csrc"
objs=""
for f in core stft mel_gemm griffinlim pianoroll resample; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $f.cu -o "$out/$f.o" &
  objs="$objs $out/$f.o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$out/libmst_b200.so" $objs -cudart static
rm -f $objs
echo "$out/libmst_b200.so"
