python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
python tools/latency_stage_probe.py 2>&1 | tail -3
python tools/stage_probe.py C1 512 2>&1 | tail -2
python tools/stage_probe.py C1 8 2>&1 | tail -2
