python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; tail -2 gpurun_out/r02e_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_ref.json 2> gpurun_out/r02e_bench_ref.err; tail -2 gpurun_out/r02e_bench_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/r02e_bench_n1.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ceiling',d['e2e']['copy_ceiling']['value'],'frac',d['e2e']['frac_of_copy_ceiling'])
print(d['roofline']['stages_ms_per_step'], d['roofline']['frame'])
print('cpu',d['cpu_baseline'])
print('knn',d['knn']['value'],d['knn']['check'],d['knn']['accepted_matches'])
print({k:(v['value'],v['roofline']['frame']['frac'],v.get('cpu_baseline',{}).get('value')) for k,v in d['configs'].items()})
print('latency',d['latency'])
r=json.loads(open('gpurun_out/r02e_bench_ref.json').read().strip().split('\n')[-1]); print('ref',r['value'],r['cpu_baseline'])
"
