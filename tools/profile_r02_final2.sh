# second set of final captures: arming / pair-search kernels at a representative harvest, launch list of the default bench
set -x
ncu --set full --import-source on --clock-control none -k regex:"ns_arm|bp_pairs2|bp_ex2|bp_pos_count|bp_stream_harvest" -s 120 -c 5 -o gpurun_out/r3e_harvest_kernels python tools/stream_bench.py --eps 0.49 --frames 8192 --harvest 0 > gpurun_out/r3e_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/r3e_launches_bench.csv python bench.py --steps 1 --warmup 1 --frames-per-graph 2048 --workloads none --no-cpu-baseline > gpurun_out/r3e_ncu2.log 2>&1
