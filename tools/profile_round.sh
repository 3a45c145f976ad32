#!/bin/bash
# tools/profile_round.sh TAG -- run on the GPU box (gpurun): the ncu evidence of one round, written to gpurun_out/.
#   1. launch list of a short bench run (gpu__time_duration.sum per launch; clocks not locked)
#   2. ncu --set full of ONE 64-picture launch of each chain kernel (with source correlation)
# A number printed by a run under ncu is never a bench value; the bench line itself comes from a plain run.
TAG=${1:-r2}
OUT=gpurun_out
ARGS="--no-cpu-baseline --no-stress --no-decoder --e2e-steps 1 --steps 2 --warmup 1"
python bench.py $ARGS > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $ARGS > $OUT/${TAG}_ncu_launches.log 2>&1
for K in k_alf k_dbf_sao; do
  ncu --target-processes application-only --set full --clock-control none --import-source on -k regex:$K --launch-skip 1 --launch-count 1 -f -o $OUT/${TAG}_$K python bench.py $ARGS > $OUT/${TAG}_ncu_$K.log 2>&1
done
ls -la $OUT/${TAG}_*
