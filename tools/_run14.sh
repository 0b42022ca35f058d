python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02c_bench_n8.json 2> gpurun_out/r02c_bench_n8.err
tail -3 gpurun_out/r02c_bench_n8.err
nvidia-smi topo -m > gpurun_out/r02c_topo.txt 2>&1
lscpu | head -20 > gpurun_out/r02c_lscpu.txt
python -c "
import json
d=json.loads(open('gpurun_out/r02c_bench_n8.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ceiling',d['e2e']['copy_ceiling'],'frac',d['e2e']['frac_of_copy_ceiling'])
print('knn',d['knn']['value'],d['knn']['check'],d['knn']['accepted_matches'])
print({k:(v['value'],v['roofline']['frame']['frac']) for k,v in d['configs'].items()})
print(d['config']['cpu_affinity'])
"
