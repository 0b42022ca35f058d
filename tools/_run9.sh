python -m pytest tests/test_cpp_dropin.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
./tests/cpp/_build/knn_sharded_test 2 3000 200003 2>&1 | tail -2
./tests/cpp/_build/knn_sharded_test 2 100000 2000000 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err
tail -2 gpurun_out/r02b_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02b_bench_n2.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ceiling',d['e2e']['copy_ceiling']['value'],'frac',d['e2e']['frac_of_copy_ceiling'])
print('knn',d['knn']['value'],d['knn']['check'],d['knn']['accepted_matches'])
print({k:(v['value'],v['roofline']['frame']['frac']) for k,v in d['configs'].items()})
"
