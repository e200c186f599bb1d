#!/bin/bash
# Timing experiments: builds fl_scaling_sc_ldpc_b200/libscldpc_v<N>.so with bp_node_kernels.cu compiled at -DNS_VARIANT=<N>
# (see the top of bp_node_kernels.cu), or libscldpc_np.so with -DNS_PREFETCH=0 for the argument "np".
# Select one with SCLDPC_LIB=...  Usage: tools/build_variants.sh 1 16 np ...
set -e
cd "$(dirname "$0")/../fl_scaling_sc_ldpc_b200/csrc"
make -s -j8
for v in "$@"; do
  if [ "$v" = np ]; then def="-DNS_PREFETCH=0"; name=np; else def="-DNS_VARIANT=$v"; name=v$v; fi
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include $def -c bp_node_kernels.cu -o /tmp/bp_node_kernels_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libscldpc_$name.so capi.o bp_kernels.o bp_wave_kernels.o /tmp/bp_node_kernels_$name.o bp_window_node_kernels.o ss_kernels.o traj_kernels.o corr_kernels.o graph_kernels.o peel_kernels.o
  echo built libscldpc_$name.so
done
