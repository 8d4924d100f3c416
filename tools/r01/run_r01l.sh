timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_r01l.txt; cat gpurun_out/pytest_r01l.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default_l.json 2> gpurun_out/bench_default_l.err; tail -c 200 gpurun_out/bench_default_l.err
for m in wilk randu goe s2; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_l.json 2> gpurun_out/bench_${m}16k_l.err; done
timeout 300 python bench.py --steps 2 --warmup 3 --matrix goe --size 32768 --no-cpu-baseline > gpurun_out/bench_goe32k_l.json 2> gpurun_out/bench_goe32k_l.err
timeout 300 python tools/resid_variants.py > gpurun_out/resid_variants_l.jsonl 2> gpurun_out/resid_variants_l.err
timeout 300 python tools/select_bench.py --sizes 4096,16384,32768,65536 --ks 1,16,64 > gpurun_out/select_bench_l.jsonl 2> gpurun_out/select_bench_l.err
timeout 300 python tools/select_bench.py --sizes 4096 --matrix s1 --ks 1,16 >> gpurun_out/select_bench_l.jsonl 2>> gpurun_out/select_bench_l.err
python - <<'PY'
import json
for f in ['default','wilk16k','randu16k','goe16k','s216k','goe32k']:
    try:
        j=json.loads(open('gpurun_out/bench_%s_l.json'%f).read().strip().splitlines()[-1])
        print(f, round(j['value']*1e3,4),'ms e2e',round(j['e2e']['value']*1e3,4), {k:round(v,3) for k,v in j['phase_ms'].items()}, j['roofline']['kernel'], round(j['roofline']['frac'],3),'res',j['check']['max_residual'],'orth',j['check'].get('orthogonality_max_abs'))
    except Exception as e: print(f,'ERR',e)
for l in open('gpurun_out/select_bench_l.jsonl'):
    j=json.loads(l); print(j['matrix'],j['n'],j['K'],round(j['device_s']*1e3,3),round(j['apply_s']*1e3,3),round(j['gpairs_per_s']))
PY
