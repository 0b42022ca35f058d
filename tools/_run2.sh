set -x
python tools/stage_probe.py C1 512 > gpurun_out/r02_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02a_launches.csv python tools/stage_probe.py C1 512 > gpurun_out/r02a_ncu_launches.log 2>&1
python tools/stage_probe.py C1 512 > gpurun_out/r02_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_fast_cells2|k_level_strip|k_pyramid_strip' -s 18 -c 6 -o gpurun_out/r02a_strip python tools/stage_probe.py C1 512 > gpurun_out/r02a_ncu_strip.log 2>&1
tail -3 gpurun_out/r02a_ncu_strip.log
