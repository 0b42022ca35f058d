python tools/stage_probe.py C4 96 2>&1 | tail -2
python tools/stage_probe.py C4 148 2>&1 | tail -2
python tools/stage_probe.py C2 256 2>&1 | tail -2
python tools/stage_probe.py C3 512 2>&1 | tail -2
