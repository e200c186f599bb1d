# A/B of the lane compaction in the tail of the frame streams (SCLDPC_COMPACT) on the headline and on the 1024-frames-per-graph workload
for c in 0 1; do
  echo "== SCLDPC_COMPACT=$c"
  SCLDPC_COMPACT=$c python bench.py --steps 4 --warmup 2 --no-cpu-baseline --workloads bp_full_fpg1024 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'] / 1e13, 4), 'e13 edge-updates/s', round(d['frames_per_s']), 'frames/s', d['gpu_launches'], 'launches; fpg1024:', json.dumps(d['workloads']['bp_full_fpg1024'])[:400])"
done
