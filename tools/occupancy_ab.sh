# A/B of ns_iter_kernel builds at different resident blocks per SM: libscldpc_nb<blocks>.so (built in a scratch copy of csrc/)
for lib in ${LIBS:-libscldpc.so libscldpc_nb3.so libscldpc_nb2.so libscldpc.so}; do
  echo "== $lib"
  SCLDPC_LIB=$PWD/fl_scaling_sc_ldpc_b200/$lib python tools/stream_bench.py --eps 0.49 --frames 4096 --harvest 0 | tail -1
done
