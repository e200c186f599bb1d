# A/B of the adaptive harvest period of the frame streams: policy (0 = fastest live graph, 1 = slowest) x constant
for cfg in "1 60" "0 35" "0 60" "1 35" "0 25" "0 45"; do
  set -- $cfg
  echo "== SCLDPC_HARVEST_POLICY=$1 SCLDPC_HARVEST_C10=$2"
  SCLDPC_HARVEST_POLICY=$1 SCLDPC_HARVEST_C10=$2 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --workloads none 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'] / 1e13, 4), 'e13 edge-updates/s', round(d['frames_per_s']), 'frames/s', d['gpu_launches'], 'launches')"
done
