timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --workload config3 --steps 5 --warmup 3 --e2e-steps 3 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.log
grep "e2e wall\|cpu baseline" gpurun_out/bench_config3.log
python bench.py --workload config4 --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.log
grep "e2e wall\|cpu baseline" gpurun_out/bench_config4.log
python bench.py --workload config2 --steps 10 --warmup 3 --e2e-steps 5 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.log
for f in gpurun_out/bench_config3.json gpurun_out/bench_config4.json gpurun_out/bench_config2.json; do python -c "
import json,sys
d=json.load(open('$f'))
print('$f', round(d['value']), round(d['e2e']['value']), d.get('config',{}).get('stage_ms'), d.get('cpu_baseline',{}).get('value'), d['roofline']['frac'], d['roofline']['kernel'])
"; done
