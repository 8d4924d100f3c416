#!/bin/bash
# r02 call G (8 GPUs): multi-rank parity test on 8 ranks, a 6-rank run (any rank count), CLI on 8 GPUs with -c and -v,
# bench at N=8 with all BASELINE configurations as extra keys (n=32768 target, n=65536 configs[4] incl. the accurate-rule check).
O=gpurun_out/r02; mkdir -p $O
nvidia-smi topo -m > $O/g_topo.txt 2>&1
timeout 900 python -m pytest "tests/test_gpu_multi_rank.py::test_sharded_solve_over_nccl[8]" -m gpu -x -q > $O/pytest_g.txt 2>&1; echo "pytest rc $?" >> $O/pytest_g.txt; tail -4 $O/pytest_g.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 6 --master-addr 127.0.0.1 --master-port 29531 tests/gpu_multi_rank_worker.py goe_n4096_p8 s1_n1000_p8 wilk > $O/world6.txt 2>&1; echo "world6 rc $?" >> $O/world6.txt; grep -a "CASE\|OK\|rc" $O/world6.txt | tail -6
timeout 300 ./cuppens -p 8 -g 8 -s 1 -n 4096 -e -c -v $O/cli_g8.bin $O/cli_g8.txt > $O/cli_g8.log 2>&1; echo "cli rc $?" >> $O/cli_g8.log; grep -a "Orthogonality\|finished\|rc" $O/cli_g8.log
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
import symmetric_eigenvalue_b200 as se
r, l, V = se.read_eigenvector_file('gpurun_out/r02/cli_g8.bin')
g = np.load('tests/golden/s1_n4096_p8.npz')
print('cli -g 8: file', V.shape, 'orth', np.abs(V.T @ V - np.eye(V.shape[1])).max(), 'lam vs reference', np.abs(l - g['lam']).max())
import os; os.unlink('gpurun_out/r02/cli_g8.bin')
PY
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_g_g8.json 2> $O/bench_g_g8.err; echo "bench rc $?" >> $O/bench_g_g8.err
tail -2 $O/bench_g_g8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_g_g8.json").read().strip().splitlines()[-1])
print("N=8", d["value"], d["check"]["parity"], d["check"]["parity_all_configs"], d["comm_backend"], {k: round(v, 3) for k, v in d["phase_ms"].items()}, d.get("same_workload_1gpu"))
for k, v in d["other_configs"].items():
    print("  ", k, v["value"], v["check"]["parity"], v.get("same_workload_1gpu"), {a: round(b, 3) for a, b in v["phase_ms"].items()}, v["roofline"]["achieved"])
    if "accurate_rule" in v: print("     accurate rule:", v["accurate_rule"]["value"], {a: b for a, b in v["accurate_rule"]["check"].items() if a in ("parity", "lambda_max_abs_diff_vs_lapack", "lambda_tol", "max_residual", "orthogonality_sampled_max_abs")})
PY
