#!/bin/bash
# r02 call D (4 GPUs): full GPU suite (multi-rank worlds 2, 3, 4), CLI on 3 GPUs with -v and -eFILE, bench at N=4.
O=gpurun_out/r02; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest_d.txt 2>&1; echo "pytest rc $?" >> $O/pytest_d.txt; tail -4 $O/pytest_d.txt
timeout 120 ./cuppens -p 8 -g 3 -s 1 -n 2048 -e -v $O/cli_g3.bin $O/cli_g3.txt > $O/cli_g3.log 2>&1; echo "cli rc $?" >> $O/cli_g3.log; tail -2 $O/cli_g3.log
timeout 120 ./cuppens -p 8 -g 1 -s 1 -n 2048 -e -v $O/cli_g1.bin $O/cli_g1.txt > $O/cli_g1.log 2>&1; echo "cli rc $?" >> $O/cli_g1.log
printf "1\n7\n2048\n100\n" > $O/ev.txt
timeout 120 ./cuppens -p 8 -g 3 -s 1 -n 2048 -e$O/ev.txt -v $O/cli_g3_sel.bin $O/cli_g3_sel.txt > $O/cli_g3_sel.log 2>&1; echo "cli rc $?" >> $O/cli_g3_sel.log; tail -2 $O/cli_g3_sel.log
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
import symmetric_eigenvalue_b200 as se
O = 'gpurun_out/r02/'
a = np.loadtxt(O + 'cli_g1.txt'); b = np.loadtxt(O + 'cli_g3.txt')
print('CLI g1 vs g3: lam diff', np.abs(a[:, 0] - b[:, 0]).max(), 'resid diff', np.abs(a[:, 1] - b[:, 1]).max())
r1, l1, V1 = se.read_eigenvector_file(O + 'cli_g1.bin'); r3, l3, V3 = se.read_eigenvector_file(O + 'cli_g3.bin')
d = np.minimum(np.abs(V1 - V3).max(axis=0), np.abs(V1 + V3).max(axis=0)).max()
print('eigenvector files g1 vs g3: shape', V3.shape, 'max diff', d, 'orth', np.abs(V3.T @ V3 - np.eye(V3.shape[1])).max())
rs, ls, Vs = se.read_eigenvector_file(O + 'cli_g3_sel.bin')
sel = rs.astype(int)
ds = np.minimum(np.abs(Vs - V1[:, sel]).max(axis=0), np.abs(Vs + V1[:, sel]).max(axis=0)).max()
print('selected on 3 GPUs: ranks', sel.tolist(), 'max diff vs full', ds)
rows = [l.split() for l in open(O + 'cli_g3_sel.txt')]
print('lines with residual', [i + 1 for i, x in enumerate(rows) if len(x) == 2])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_d_g4.json 2> $O/bench_d_g4.err; echo "bench rc $?" >> $O/bench_d_g4.err
tail -2 $O/bench_d_g4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_d_g4.json").read().strip().splitlines()[-1])
print("N=4", d["value"], d["check"]["parity"], d["comm_backend"], {k: round(v, 3) for k, v in d["phase_ms"].items()}, d.get("same_workload_1gpu"))
for k, v in d["other_configs"].items(): print("  ", k, v["value"], v["check"]["parity"], v.get("same_workload_1gpu"))
PY
