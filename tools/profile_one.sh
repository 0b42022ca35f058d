#!/bin/bash
# One ncu capture per gpurun call: runs the command plain first (must exit 0), then once under ncu.
#   tools/profile_one.sh <round> launches                      -> gpurun_out/<round>_launches.csv (gpu__time_duration of a window of the bench)
#   tools/profile_one.sh <round> <kernel> <skip> [probe|bench] -> gpurun_out/<round>_<kernel>.ncu-rep (--set full, one launch)
set -u
R=$1; K=$2; SKIP=${3:-0}; WHAT=${4:-bench}
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --batch 256 --skip-cpu --knn-steps 1 --knn-queries 131072 --knn-train-per-gpu 262144"
PROBE="python tools/guided_probe.py --quick"
CMD=$BENCH; [ "$WHAT" = probe ] && CMD=$PROBE
$CMD > gpurun_out/${R}_plain.out 2> gpurun_out/${R}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.err; exit 1; }
if [ "$K" = launches ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
else
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/${R}_${K} $CMD > gpurun_out/${R}_ncu_${K}.log 2>&1
fi
echo "ncu rc=$?"
