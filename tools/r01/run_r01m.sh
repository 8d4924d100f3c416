TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --matrix goe --size 16384 > gpurun_out/scale4_goe_16384_g2.json 2> gpurun_out/scale4_goe_16384_g2.err
timeout 200 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 --matrix s1 --size 4096 --no-single-gpu-compare > gpurun_out/scale4_s1_4096_g2.json 2> gpurun_out/scale4_s1_4096_g2.err
timeout 200 ./cuppens -p 8 -g 2 -s 2 -n 2048 -e gpurun_out/cli_g2_out.txt > gpurun_out/cli_g2_v2.txt 2>&1; tail -4 gpurun_out/cli_g2_v2.txt; head -3 gpurun_out/cli_g2_out.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale4_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j['n_gpus'], round(j['value']*1e3,3),'ms', {k:round(v,2) for k,v in j['phase_ms'].items()}, 'res', j['check']['max_residual'], j.get('same_workload_1gpu'))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
