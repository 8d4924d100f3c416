#!/bin/bash
# r02 call H5 (1 GPU): dense back-transformation with the parallel Gram / dlarft, U kernel with fewer column chunks per row tile.
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_h5.txt 2>&1; echo "pytest rc $?" >> $O/pytest_h5.txt; tail -4 $O/pytest_h5.txt
timeout 600 python tools/dense_bench.py 1024 4096 > $O/dense_bench_h5.txt 2>&1; tail -4 $O/dense_bench_h5.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_h5.json 2> $O/bench_h5.err; echo "bench rc $?" >> $O/bench_h5.err; tail -1 $O/bench_h5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_h5.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["check"]["parity_all_configs"], d["roofline"]["achieved"], d["launches_per_step"], {k: round(v, 3) for k, v in d["phase_ms"].items()})
for k, v in d["other_configs"].items(): print("  ", k, v["value"], v["launches_per_step"], v["check"]["parity"], {a: round(b, 3) for a, b in v["phase_ms"].items()})
PY
