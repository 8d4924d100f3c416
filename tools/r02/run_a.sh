#!/bin/bash
# r02 call A (1 GPU): tensor-map TMA probe (plain, then memcheck only if the plain run passed), GPU test suite,
# smoke, default bench line, launch list of the headline workload.
O=gpurun_out/r02; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > $O/a_smi.txt 2>&1
( ./tests/probe/tma_probe; echo "exit code $?" ) > $O/tma_probe_plain.txt 2>&1
( ./tests/probe/tma_min 0 0; echo "exit code $?" ) > $O/tma_min_plain.txt 2>&1
if grep -q "^0 case(s) failed" $O/tma_probe_plain.txt; then
  ( timeout 300 compute-sanitizer --tool memcheck ./tests/probe/tma_probe; echo "exit code $?" ) > $O/tma_probe_memcheck.txt 2>&1
fi
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_a.txt 2>&1; echo "pytest rc $?" >> $O/pytest_a.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_a.txt 2>&1; echo "smoke rc $?" >> $O/smoke_a.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_a.json 2> $O/bench_a.err; echo "bench rc $?" >> $O/bench_a.err
tail -3 $O/pytest_a.txt; cat $O/tma_probe_plain.txt; tail -2 $O/smoke_a.txt; tail -2 $O/bench_a.err
