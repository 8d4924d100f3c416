#!/bin/bash
# r02 call W (8 GPUs, short): the final code at N=8 on the headline workload only (parity verdict + time per step)
O=gpurun_out/r02; mkdir -p $O
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --steps 6 --warmup 3 --no-extras --no-cpu-baseline --select 0 > $O/bench_w_g8.json 2> $O/bench_w_g8.err; echo "bench rc $?" >> $O/bench_w_g8.err
grep -v "^\*\*\*\|OMP_NUM" $O/bench_w_g8.err | tail -3 | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_w_g8.json").read().strip().splitlines()[-1])
print("N=8", d["value"], d["check"]["parity"], d["comm_backend"], {k: round(v, 3) for k, v in d["phase_ms"].items()}, d.get("same_workload_1gpu"), "e2e", d["e2e"]["value"])
PY
