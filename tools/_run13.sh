python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
for k in "" "ORBX_CELLS_PER_WARP=1"; do env $k python tools/stage_probe.py C1 512 2>&1 | tail -2; done
python tools/stage_probe.py C4 32 2>&1 | tail -2
