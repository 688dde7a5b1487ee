#!/bin/sh
# Instrumented build of the row-streaming fused kernel (per-phase clock64 totals) for tools/phase_clocks.py:
#   tools/build_clocks.sh && MPCG_B200_LIB=tools/libmpcg_b200_clocks.so python tools/phase_clocks.py
set -e
cd /root/repo/wav2vec-heart-sounds_b200/csrc
d=/tmp/mpcg_clocks_build; rm -rf $d; mkdir -p $d
for f in stream.cu stream_inst_t33_16.cu stream_inst_t8.cu; do
  nvcc -DMPCG_FZ_PHASE_CLOCKS=1 -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -c $f -o $d/${f%.cu}.o &
done
wait
objs=$(ls build/*.o | grep -v "build/stream.o\|stream_inst_t33_16\|stream_inst_t8")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/libmpcg_b200_clocks.so $objs $d/*.o -lcudart
echo built tools/libmpcg_b200_clocks.so
