# A/B of the harvest period constant after the tile-based arming kernel (cheaper harvests allow shorter periods)
for c in 60 35 20 12; do
  echo "== SCLDPC_HARVEST_C10=$c"
  SCLDPC_HARVEST_C10=$c python bench.py --steps 4 --warmup 2 --no-cpu-baseline --workloads none 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'] / 1e13, 4), 'e13 edge-updates/s', round(d['frames_per_s']), 'frames/s', d['gpu_launches'], 'launches')"
done
