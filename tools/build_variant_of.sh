#!/bin/sh
# Build a variant of ONE translation unit next to the product library, for A/B timing:
#   tools/build_variant_of.sh NAME aug_chain.cu "-DMPCG_AC_P1_WEIGHTS=1"  ->  tools/libmpcg_b200_NAME.so  (MPCG_B200_LIB=...)
set -e
name=$1; src=$2; flags=$3
cd /root/repo/wav2vec-heart-sounds_b200/csrc
d=/tmp/mpcg_variant_$name; rm -rf $d; mkdir -p $d
nvcc $flags -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -c $src -o $d/${src%.cu}.o 2> $d/log || { grep -i error $d/log | head; exit 1; }
grep -o "[0-9]* bytes spill stores, [0-9]* bytes spill loads" $d/log | sort | uniq -c | tail -3
objs=$(ls build/*.o | grep -v "build/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/libmpcg_b200_$name.so $objs $d/${src%.cu}.o -lcudart
echo built tools/libmpcg_b200_$name.so
