timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/pytest_r01t.txt; cat gpurun_out/pytest_r01t.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_t.json 2> gpurun_out/bench_default_t.err; tail -c 200 gpurun_out/bench_default_t.err
timeout 300 python bench.py --steps 3 --warmup 3 --matrix goe --size 16384 --no-cpu-baseline --select 0 > gpurun_out/bench_goe16k_t.json 2> gpurun_out/bench_goe16k_t.err
python - <<'PY'
import json
for f in ['default','goe16k']:
    try:
        j=json.loads(open('gpurun_out/bench_%s_t.json'%f).read().strip().splitlines()[-1])
        print(f, round(j['value']*1e3,4),'ms e2e',round(j['e2e']['value']*1e3,4), {k:round(v,3) for k,v in j['phase_ms'].items()}, j['gpu_launches'], j['check'])
    except Exception as e: print(f,'ERR',e)
PY
