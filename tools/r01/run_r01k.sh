timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_r01k.txt; cat gpurun_out/pytest_r01k.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_k.json 2> gpurun_out/bench_default_k.err; tail -c 200 gpurun_out/bench_default_k.err
for m in goe; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_k.json 2> gpurun_out/bench_${m}16k_k.err; done
timeout 300 python tools/select_bench.py --sizes 4096,16384,32768,65536 --ks 1,16,64 > gpurun_out/select_bench_k.jsonl 2> gpurun_out/select_bench_k.err
timeout 300 python tools/select_bench.py --sizes 4096 --matrix s1 --ks 1,16 >> gpurun_out/select_bench_k.jsonl 2>> gpurun_out/select_bench_k.err
python - <<'PY'
import json
for f in ['default','goe16k']:
    j=json.loads(open('gpurun_out/bench_%s_k.json'%f).read().strip().splitlines()[-1])
    print(f, round(j['value']*1e3,4),'ms e2e',round(j['e2e']['value']*1e3,4), {k:round(v,3) for k,v in j['phase_ms'].items()}, 'sel16',round(j['selected_mode']['device_s_per_solve']*1e3,3))
for l in open('gpurun_out/select_bench_k.jsonl'):
    j=json.loads(l); print(j['matrix'],j['n'],j['K'],round(j['device_s']*1e3,3),round(j['apply_s']*1e3,3),round(j['gpairs_per_s']))
PY
