#!/bin/bash
# usage: tools/prof_fwd.sh <tag> <kernel-regex> [kbench args]   (run on the GPU box through gpurun)
TAG=$1; KR=$2; shift 2
CMD="python tools/kbench.py --iters 5 $*"
$CMD && ncu --set full --clock-control none --import-source on -k regex:$KR -s 3 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
