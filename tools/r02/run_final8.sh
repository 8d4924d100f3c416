#!/bin/bash
# r02 final 8-GPU call: multi-rank parity test on 8 ranks + NCCL-fallback test, smoke(), bench at N=8 (all configurations).
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest "tests/test_gpu_multi_rank.py::test_sharded_solve_over_nccl[8]" tests/test_gpu_multi_rank.py::test_sharded_solve_collective_fallback -m gpu -x -q > $O/pytest_final8.txt 2>&1; echo "pytest rc $?" >> $O/pytest_final8.txt; tail -3 $O/pytest_final8.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final8.txt 2>&1; echo "smoke rc $?" >> $O/smoke_final8.txt; tail -2 $O/smoke_final8.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_final_g8.json 2> $O/bench_final_g8.err; echo "bench rc $?" >> $O/bench_final_g8.err
grep -v "^\*\*\*\|OMP_NUM" $O/bench_final_g8.err | tail -3 | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_final_g8.json").read().strip().splitlines()[-1])
print("N=8", d["value"], d["check"]["parity"], d["check"]["parity_all_configs"], d["comm_backend"], {k: round(v, 3) for k, v in d["phase_ms"].items()}, d.get("same_workload_1gpu"), "e2e", d["e2e"]["value"])
for k, v in d["other_configs"].items():
    print("  ", k, v["value"], v["check"]["parity"], v.get("same_workload_1gpu"), {a: round(b, 3) for a, b in v["phase_ms"].items()}, v["roofline"]["achieved"])
    if "accurate_rule" in v: print("     accurate rule:", v["accurate_rule"]["value"], {a: b for a, b in v["accurate_rule"]["check"].items() if a in ("parity", "lambda_max_abs_diff_vs_lapack", "lambda_tol", "max_residual", "orthogonality_sampled_max_abs")})
PY
