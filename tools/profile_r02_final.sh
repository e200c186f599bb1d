# ncu captures of round 2's final kernels (one GPU; every command has run to completion without ncu before)
set -x
ncu --set full --import-source on --clock-control none -k regex:ns_iter -s 700 -c 2 -o gpurun_out/r2y_ns_iter python tools/stream_bench.py --frames 3072 --eps 0.48 > gpurun_out/r2y_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"ns_arm|ns_compact_kernel" -s 3 -c 4 -o gpurun_out/r2y_ns_arm python tools/stream_bench.py --frames 3072 --eps 0.48 > gpurun_out/r2y_ncu2.log 2>&1
SCLDPC_PEEL_SLOTS=5328 ncu --set full --import-source on --clock-control none -k regex:peel_trajectory -c 1 -o gpurun_out/r2y_peel_M10000 python tools/peel_slots_sweep.py --one 5328 --M 10000 --graphs 16 --frames 333 > gpurun_out/r2y_ncu3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:peel_trajectory -c 1 -o gpurun_out/r2y_peel_M1000 python tools/peel_slots_sweep.py --one 5328 --M 1000 --graphs 16 --frames 333 > gpurun_out/r2y_ncu4.log 2>&1
ls -la gpurun_out/r2y_*
