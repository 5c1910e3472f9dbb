#!/bin/bash
# Round-end measurement on ONE B200 (run through gpurun): bench line, sweep, ncu launch list + --set full captures.
# Everything lands under gpurun_out/ with the given tag; profiles/summarize.py turns the captures into profiles/*.md.
TAG=${1:-r2_final}
O=gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err
python tools/sweep.py --out $O/${TAG}_sweep.md > $O/${TAG}_sweep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-train --no-cpu --no-extra > $O/${TAG}_ncu_l.log 2>&1
for W in cfg2 cfg2_bf16 cfg1 cfg3; do
  PROF_WHAT=$W ncu --set full --clock-control none --import-source on -k regex:"vq_fwd|vq_bwd" -c 2 -o $O/${TAG}_prof_$W -f \
      python tools/prof_r2.py > $O/${TAG}_ncu_$W.log 2>&1
done
ls $O/${TAG}_*
