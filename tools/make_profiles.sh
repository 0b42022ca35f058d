#!/bin/bash
# Run on a B200 (under gpurun): the launch list of a short bench.py run, then ONE --set full capture that holds one whole device-resident
# step (7 pyramid levels, blur, dense FAST bound, cells, quadtree, describe) plus the kNN kernel. Outputs go to gpurun_out/;
# tools/summarize_profiles.py turns them into the text summaries under profiles/. usage: tools/make_profiles.sh r02
set -u
R=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --batch 256 --skip-cpu --skip-stereo --skip-configs --skip-guided --knn-steps 1 --knn-queries 131072 --knn-train-per-gpu 262144"
$CMD > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.err; exit 1; }
# the window starts inside the timed device-resident steps (3 warm-up + 3 timed steps of 13 launches + the D2D repack each)
ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 120 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_pyramid_strip|k_level_strip|k_fast_cells2|k_quadtree|k_orient_describe2' -s 36 -c 12 -f -o gpurun_out/${R}_step $CMD > gpurun_out/${R}_ncu_step.log 2>&1
echo "step capture rc=$?"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_knn2_partial -c 1 -f -o gpurun_out/${R}_k_knn2_partial $CMD > gpurun_out/${R}_ncu_knn.log 2>&1
echo "knn capture rc=$?"
