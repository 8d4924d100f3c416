#!/bin/bash
# r02 call F0 (1 GPU): suite incl. the dense front end, bench after the fast-reciprocal / work-list changes, launch list.
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_f0.txt 2>&1; echo "pytest rc $?" >> $O/pytest_f0.txt; tail -12 $O/pytest_f0.txt
timeout 600 python tools/dense_bench.py 1024 4096 8192 > $O/dense_bench.txt 2>&1; cat $O/dense_bench.txt | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_f0.json 2> $O/bench_f0.err; echo "bench rc $?" >> $O/bench_f0.err; tail -1 $O/bench_f0.err
python tools/profile_step.py --size 16384 --matrix goe > $O/prof_plain_f0.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 103 -c 110 --csv --log-file $O/launches_goe16k_f0.csv python tools/profile_step.py --size 16384 --matrix goe > $O/ncu_launch_f0.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_f0.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["check"]["parity"], d["roofline"]["achieved"], d["roofline"]["frac"], {k: round(v, 3) for k, v in d["phase_ms"].items()})
for k, v in d["other_configs"].items(): print("  ", k, v["value"], v["launches_per_step"], v["check"]["parity"], {a: round(b, 3) for a, b in v["phase_ms"].items()})
print("  eig-only", d["eigenvalues_only"], "select", d["selected_mode"]["device_s_per_solve"])
PY
