python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -2
python tools/latency_probe2.py 2>&1 | tail -2
ORBX_NO_GRAPH=1 python tools/latency_probe2.py 2>&1 | tail -2
