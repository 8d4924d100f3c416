#!/bin/bash
# r02 call B (2 GPUs): NCCL / peer-memory multi-rank parity test, CLI with -g 2, bench at N=2 (peer memory) and with
# CUPPEN_P2P=0 (NCCL collectives) for comparison.
O=gpurun_out/r02; mkdir -p $O
nvidia-smi topo -m > $O/b_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi_rank.py -m gpu -x -q > $O/pytest_b.txt 2>&1; echo "pytest rc $?" >> $O/pytest_b.txt
tail -5 $O/pytest_b.txt
timeout 120 ./cuppens -p 8 -g 2 -s 1 -n 2048 -e $O/cli_g2.txt > $O/cli_g2.log 2>&1; echo "cli rc $?" >> $O/cli_g2.log; tail -2 $O/cli_g2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_b_g2.json 2> $O/bench_b_g2.err; echo "bench rc $?" >> $O/bench_b_g2.err
tail -2 $O/bench_b_g2.err
CUPPEN_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > $O/bench_b_g2_nccl.json 2> $O/bench_b_g2_nccl.err; echo "bench rc $?" >> $O/bench_b_g2_nccl.err
tail -2 $O/bench_b_g2_nccl.err
python - <<'PY'
import json
for f in ("bench_b_g2.json","bench_b_g2_nccl.json"):
    try:
        d=json.loads(open("gpurun_out/r02/"+f).read().strip().splitlines()[-1])
        print(f, d["value"], d["check"]["parity"], d["phase_ms"], d.get("same_workload_1gpu"))
    except Exception as e: print(f, "unreadable", e)
PY
