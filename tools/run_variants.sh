# A/B of libscldpc builds on the stream micro-benchmark (one graph, eps = 0.48)
for lib in "$@"; do
  echo "== $lib"; SCLDPC_LIB=$PWD/fl_scaling_sc_ldpc_b200/$lib timeout 120 python tools/stream_bench.py --frames 3072 --eps 0.48 2>&1 | tail -1
  SCLDPC_LIB=$PWD/fl_scaling_sc_ldpc_b200/$lib timeout 120 python tools/stream_bench.py --frames 3072 --eps 0.49 2>&1 | tail -1
done
