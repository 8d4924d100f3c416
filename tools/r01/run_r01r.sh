timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/pytest_r01r.txt; cat gpurun_out/pytest_r01r.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_r.json 2> gpurun_out/bench_default_r.err; tail -c 200 gpurun_out/bench_default_r.err
for m in wilk goe; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_r.json 2> gpurun_out/bench_${m}16k_r.err; done
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import json
for f in ['default','wilk16k','goe16k']:
    try:
        j=json.loads(open('gpurun_out/bench_%s_r.json'%f).read().strip().splitlines()[-1])
        print(f, round(j['value']*1e3,4),'ms e2e',round(j['e2e']['value']*1e3,4), {k:round(v,3) for k,v in j['phase_ms'].items()}, j['roofline']['kernel'], round(j['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
