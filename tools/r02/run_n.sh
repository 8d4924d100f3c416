#!/bin/bash
# r02 call N (2 GPUs): multi-rank parity on two ranks (peer memory) with the row supports, and the NCCL-collective fallback
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest "tests/test_gpu_multi_rank.py::test_sharded_solve_over_nccl[2]" tests/test_gpu_multi_rank.py::test_sharded_solve_collective_fallback -m gpu -x -q > $O/pytest_n.txt 2>&1; echo "pytest rc $?" >> $O/pytest_n.txt; tail -5 $O/pytest_n.txt
