python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/stage_probe.py C1 512 2>&1 | tail -2
ORBX_NO_DIRECT=1 python tools/stage_probe.py C1 512 2>&1 | tail -2
python tools/stage_probe.py C3 512 2>&1 | tail -2
