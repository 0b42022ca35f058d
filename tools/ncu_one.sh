#!/bin/bash
# One --set full capture of one kernel inside tools/stage_probe.py (run under gpurun). usage: tools/ncu_one.sh <kernel regex> <output name> [config] [batch] [skip]
K=$1; O=$2; C=${3:-C1}; B=${4:-128}; S=${5:-3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/$O python tools/stage_probe.py $C $B > gpurun_out/$O.log 2>&1
tail -2 gpurun_out/$O.log
