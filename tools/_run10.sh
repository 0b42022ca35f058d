python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
for k in "" "ORBX_BLUR_SIDE=1" ; do env $k python tools/stage_probe.py C1 512 2>&1 | tail -2; done
python tools/stage_probe.py C1 512 > gpurun_out/r02_plain10.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_fast_cells2|k_level_strip' -s 6 -c 3 -o gpurun_out/r02d_fast python tools/stage_probe.py C1 512 > gpurun_out/r02d_ncu.log 2>&1
tail -2 gpurun_out/r02d_ncu.log
