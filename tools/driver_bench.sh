# Throughput of the drop-in driver a user of the reference runs (bp_cli bp_lim_iter: rounds of batches, records all-gathered and the
# sequential early stop replayed on every rank): wall time of two epsilon points of FRAMES frames at the BASELINE size, 1 rank and N ranks.
# usage: bash tools/driver_bench.sh [N]
N=${1:-1}
ARGS="bp_lim_iter 0 0 0 1000 --M 5000 --points 2 --eps-ini 0.48 --eps-delta 0.01 --max-frames ${FRAMES:-16384} --min-frame-err 1000000 --frames-per-graph ${FPG:-64} --graphs-per-batch ${GPB:-16} --seed 7"
for n in 1 $N; do
  out=$(mktemp -d)
  t0=$(date +%s.%N)
  if [ "$n" = 1 ]; then python -m fl_scaling_sc_ldpc_b200.bp_cli $ARGS --outdir $out > /dev/null 2> $out/err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 -m fl_scaling_sc_ldpc_b200.bp_cli $ARGS --outdir $out > /dev/null 2> $out/err; fi
  t1=$(date +%s.%N)
  python - <<PY
import glob
f = glob.glob("$out/*.dat")[0]
rows = open(f).read().strip().splitlines()[1:]
frames = sum(int(r.split()[9]) for r in rows)
dt = $t1 - $t0
print({"ranks": $n, "frames_per_graph": ${FPG:-64}, "frames": frames, "wall_s": round(dt, 2), "frames_per_s_incl_startup": round(frames / dt, 1), "rows": rows})
PY
  [ "$N" = 1 ] && break
done
