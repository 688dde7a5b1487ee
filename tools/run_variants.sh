for v in g1h g2 g2h; do
  echo "== $v"
  MPCG_B200_LIB=tools/libmpcg_b200_$v.so timeout 100 python tools/bench_stream.py c2 2>&1 | cut -c1-120
  MPCG_B200_LIB=tools/libmpcg_b200_$v.so timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:fused_stream -s 4 -c 1 python tools/bench_stream.py c2 2>&1 | grep -E "dram__|gpu__time"
done
