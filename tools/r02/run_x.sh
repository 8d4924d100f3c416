#!/bin/bash
# r02 call X (1 GPU, the last GPU seconds of the round): the new switch-agreement test only
O=gpurun_out/r02; mkdir -p $O
timeout 70 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "switches_agree or fast_reciprocal" > $O/pytest_x.txt 2>&1; echo "pytest rc $?" >> $O/pytest_x.txt; tail -5 $O/pytest_x.txt
