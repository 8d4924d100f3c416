timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_r01p.txt; cat gpurun_out/pytest_r01p.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default_p.json 2> gpurun_out/bench_default_p.err; tail -c 200 gpurun_out/bench_default_p.err
for m in wilk goe; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_p.json 2> gpurun_out/bench_${m}16k_p.err; done
timeout 300 python tools/select_bench.py --sizes 4096,16384 --ks 16 > gpurun_out/select_bench_p.jsonl 2> gpurun_out/select_bench_p.err
timeout 300 python tools/select_bench.py --sizes 4096 --matrix s1 --ks 1,16 >> gpurun_out/select_bench_p.jsonl 2>> gpurun_out/select_bench_p.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_s1_4096_p.csv python tools/profile_step.py --size 4096 --matrix s1 > gpurun_out/ncu_s1_p.log 2>&1
python - <<'PY'
import json
for f in ['default','wilk16k','goe16k']:
    try:
        j=json.loads(open('gpurun_out/bench_%s_p.json'%f).read().strip().splitlines()[-1])
        print(f, round(j['value']*1e3,4),'ms e2e',round(j['e2e']['value']*1e3,4), {k:round(v,3) for k,v in j['phase_ms'].items()}, j['roofline']['kernel'], round(j['roofline']['frac'],3),'res',j['check']['max_residual'],'orth',j['check'].get('orthogonality_max_abs'), j['roofline'].get('note'), j['roofline'].get('traffic'))
    except Exception as e: print(f,'ERR',e)
for l in open('gpurun_out/select_bench_p.jsonl'):
    j=json.loads(l); print(j['matrix'],j['n'],j['K'],round(j['device_s']*1e3,3),round(j['apply_s']*1e3,3),round(j['gpairs_per_s']))
PY
