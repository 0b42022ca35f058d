#!/bin/bash
# Run on a B200 (under gpurun): launch list of the bench command + one full capture of the dominant kernels.
# Outputs go to gpurun_out/; tools/summarize_profiles.py turns them into the text summaries under profiles/.
set -u
R=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --batch 256 --skip-cpu --skip-stereo --knn-steps 1 --knn-queries 131072 --knn-train-per-gpu 262144"
$CMD > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.err; exit 1; }
# device-resident steps: 6 x 18 launches (+1 D2D copy each), then e2e chunks; capture a window that covers one whole step
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 140 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rc=$?"
for spec in k_fast_cells:4 k_gauss7:24 k_pyramid_resize:21 k_orient_describe:4 k_quadtree:4 k_knn2_partial:1; do
  k=${spec%%:*}; skip=${spec##*:}
  $CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${R}_${k} $CMD > gpurun_out/${R}_ncu_${k}.log 2>&1
  echo "$k rc=$?"
done
