#!/bin/bash
# r02 call G3 (8 GPUs): bench at N=8 with all BASELINE configurations as extra keys.
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_g8.json 2> $O/bench_g8.err; echo "bench rc $?" >> $O/bench_g8.err
grep -v "^\*\*\*\|OMP_NUM" $O/bench_g8.err | tail -5 | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_g8.json").read().strip().splitlines()[-1])
print("N=8", d["value"], d["check"]["parity"], d["check"]["parity_all_configs"], d["comm_backend"], {k: round(v, 3) for k, v in d["phase_ms"].items()}, d.get("same_workload_1gpu"), "e2e", d["e2e"]["value"])
for k, v in d["other_configs"].items():
    print("  ", k, v["value"], v["check"]["parity"], v.get("same_workload_1gpu"), {a: round(b, 3) for a, b in v["phase_ms"].items()}, v["roofline"]["achieved"])
    if "accurate_rule" in v: print("     accurate rule:", v["accurate_rule"]["value"], {a: b for a, b in v["accurate_rule"]["check"].items() if a in ("parity", "lambda_max_abs_diff_vs_lapack", "lambda_tol", "max_residual", "orthogonality_sampled_max_abs")})
PY
