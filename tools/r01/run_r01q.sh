timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/pytest_r01q.txt; cat gpurun_out/pytest_r01q.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default_q.json 2> gpurun_out/bench_default_q.err; tail -c 200 gpurun_out/bench_default_q.err
for m in wilk; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_q.json 2> gpurun_out/bench_${m}16k_q.err; done
python - <<'PY'
import json
for f in ['default','wilk16k']:
    try:
        j=json.loads(open('gpurun_out/bench_%s_q.json'%f).read().strip().splitlines()[-1])
        print(f, round(j['value']*1e3,4),'ms e2e',round(j['e2e']['value']*1e3,4), {k:round(v,3) for k,v in j['phase_ms'].items()}, j['roofline']['kernel'], round(j['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
