# A/B of grids sized by the graphs still decoding (SCLDPC_ALIVE_GRIDS): 4 graphs per batch, one of them decodes alone for 85 % of the step
for c in 0 1 0 1; do
  echo "== SCLDPC_ALIVE_GRIDS=$c"
  SCLDPC_ALIVE_GRIDS=$c python bench.py --steps 5 --warmup 2 --no-cpu-baseline --workloads bp_full_fpg1024 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'] / 1e13, 4), 'e13 edge-updates/s', round(d['frames_per_s']), 'frames/s', d['gpu_launches'], 'launches', d['roofline']['p10_p50_p90_ms'], 'fpg1024:', round(d['workloads']['bp_full_fpg1024']['value']/1e13, 4))"
done
