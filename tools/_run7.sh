python tools/stage_probe.py C1 512 > gpurun_out/r02_plain7.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_orient_describe2|k_quadtree' -s 4 -c 2 -o gpurun_out/r02c_desc_qt python tools/stage_probe.py C1 512 > gpurun_out/r02c_ncu.log 2>&1
tail -2 gpurun_out/r02c_ncu.log
