# A/B of libscldpc builds on the stream micro-benchmark (one graph, eps = 0.48), then an ncu capture of the iteration kernel
# in the recycling phase of the second decode (the first ~340 launches are the warm-up decode: skip 700)
for lib in "$@"; do
  echo "== $lib"; SCLDPC_LIB=$PWD/fl_scaling_sc_ldpc_b200/$lib timeout 120 python tools/stream_bench.py --frames 3072 --eps 0.48 2>&1 | tail -1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns_iter -s 700 -c 3 -o gpurun_out/r2c_ns_iter python tools/stream_bench.py --frames 3072 --eps 0.48 > gpurun_out/r2c_ncu.log 2>&1; tail -2 gpurun_out/r2c_ncu.log
