TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { N=$1; SZ=$2; ST=$3; PORT=$4; shift 4
  timeout 600 $TR --nproc-per-node $N --master-port $PORT bench.py --gpus $N --steps $ST --warmup 3 --matrix goe --size $SZ --no-single-gpu-compare "$@" > gpurun_out/scale3_goe_${SZ}_g$N.json 2> gpurun_out/scale3_goe_${SZ}_g$N.err
  tail -c 400 gpurun_out/scale3_goe_${SZ}_g$N.json | head -c 400; echo; tail -2 gpurun_out/scale3_goe_${SZ}_g$N.err; }
run 8 32768 3 29511
run 8 65536 2 29512
run 8 16384 3 29513
run 4 32768 3 29514
run 2 16384 3 29515
timeout 300 $TR --nproc-per-node 8 --master-port 29516 bench.py --gpus 8 --steps 5 --warmup 3 --matrix s1 --size 4096 --no-single-gpu-compare > gpurun_out/scale3_s1_4096_g8.json 2> gpurun_out/scale3_s1_4096_g8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale3_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j['n_gpus'], round(j['value']*1e3,3),'ms', {k:round(v,2) for k,v in j['phase_ms'].items()}, 'gemm TF', round(j['gemm_tflops_executed'],1), 'res', j['check']['max_residual'])
    except Exception as e:
        print(f, 'ERR', e)
PY
