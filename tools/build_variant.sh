#!/bin/sh
# Build a variant of the row-streaming fused kernel next to the product library, for A/B timing:
#   tools/build_variant.sh NAME "-DMPCG_SK_THREADS=1024"   ->  tools/libmpcg_b200_NAME.so   (load it with MPCG_B200_LIB=...)
set -e
name=$1; flags=$2
cd /root/repo/wav2vec-heart-sounds_b200/csrc
d=/tmp/mpcg_variant_$name; rm -rf $d; mkdir -p $d
for f in stream.cu stream_inst_t33_16.cu stream_inst_t8.cu stream_inst_t33_32.cu; do
  nvcc $flags -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -c $f -o $d/${f%.cu}.o 2> $d/${f%.cu}.log &
done
wait
for f in stream.cu stream_inst_t33_16.cu stream_inst_t8.cu stream_inst_t33_32.cu; do
  test -f $d/${f%.cu}.o || { cat $d/${f%.cu}.log | grep -i error | head -5; echo "variant build failed: $f"; exit 1; }
done
grep -h -A2 "Compiling entry function.*fused_stream" $d/*.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores, [0-9]* bytes spill loads" | paste -sd' ' || true
objs=$(ls build/*.o | grep -v "build/stream.o\|stream_inst_t33_16\|stream_inst_t8\|stream_inst_t33_32")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/libmpcg_b200_$name.so $objs $d/*.o -lcudart
echo built tools/libmpcg_b200_$name.so
