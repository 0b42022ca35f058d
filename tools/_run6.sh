python -m pytest tests/test_gpu_cos_sin_sweep.py tests/test_gpu_extract.py tests/test_cpp_dropin.py -m gpu -x -q 2>&1 | tail -12
( time python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err ) 2>&1 | tail -4
tail -3 gpurun_out/r02b_bench_n1.err
python tools/latency_stage_probe.py 2>&1 | tail -3
