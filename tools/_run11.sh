python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -2
python tools/stage_probe.py C1 512 2>&1 | tail -2
python tools/stage_probe.py C4 32 2>&1 | tail -2
