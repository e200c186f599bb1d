# A/B of peeling kernel builds (set-up loop unroll PEEL_SU x register target PEEL_MINB): libscldpc_p<SU>_<MINB>.so
for lib in libscldpc_p1_20.so libscldpc_p1_22.so libscldpc_p1_24.so libscldpc_p1_28.so libscldpc_p2_24.so libscldpc_p1_20.so; do
  echo "== $lib"
  for M in 1000 10000; do
    SCLDPC_LIB=$PWD/fl_scaling_sc_ldpc_b200/$lib python tools/peel_slots_sweep.py --one 100000 --M $M --graphs 16 --frames 1024 | tail -1
  done
done
