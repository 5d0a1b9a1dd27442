#!/bin/bash
# usage (on the GPU box, through gpurun): tools/prof_round.sh <outdir>
# The round's evidence in one call: the contract line, the reference arm, the ncu launch list of the bench command and
# `ncu --set full` captures of one forward and one backward of a general-placement step and of a translation step.
OUT=${1:-gpurun_out/prof}
mkdir -p $OUT
python bench.py > $OUT/bench.json 2> $OUT/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
CMD="python bench.py --steps 3 --warmup 3 --kernels-only"
$CMD > $OUT/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
for TH in I P; do
  K="python tools/kbench.py --iters 3 --workloads c2 --thetas $TH"
  # kbench: 3 warm-up + 3 timed forwards (3 launches each), then 3 + 3 backwards (4 launches each)
  $K > $OUT/kbench_$TH.log 2>&1 && ncu --set full --clock-control none -s 9 -c 3 -o $OUT/prof_${TH}_fwd $K > $OUT/ncu_${TH}_fwd.log 2>&1
  ncu --set full --clock-control none -s 30 -c 4 -o $OUT/prof_${TH}_bwd $K > $OUT/ncu_${TH}_bwd.log 2>&1
done
# the reports themselves are too big to travel back (64 MiB cap): summarise here, keep the text
for f in $OUT/prof_*.ncu-rep; do python profiles/summarize_ncu.py $f > ${f%.ncu-rep}.txt; done
python tools/update_traffic.py $OUT/prof_I_fwd.ncu-rep $OUT/prof_I_bwd.ncu-rep $OUT/prof_P_fwd.ncu-rep $OUT/prof_P_bwd.ncu-rep c2 > $OUT/traffic.log 2>&1
cp profiles/roofline_traffic.json $OUT/roofline_traffic.json
rm -f $OUT/prof_*.ncu-rep
tail -c 600 $OUT/bench.json; echo; cat $OUT/kbench_I.log $OUT/kbench_P.log
