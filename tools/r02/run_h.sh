#!/bin/bash
# r02 call H (1 GPU): shared-memory fused front end (m <= 512) and the faster dense symv: suite, dense bench, bench, launch list of -s 1 -n 4096.
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_h2.txt 2>&1; echo "pytest rc $?" >> $O/pytest_h2.txt; tail -12 $O/pytest_h2.txt
timeout 600 python tools/dense_bench.py 1024 4096 8192 > $O/dense_bench_h2.txt 2>&1; tail -3 $O/dense_bench_h2.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_h2.json 2> $O/bench_h2.err; echo "bench rc $?" >> $O/bench_h2.err; tail -1 $O/bench_h2.err
python tools/profile_step.py --size 4096 --matrix s1 > $O/prof_plain_h2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 61 -c 70 --csv --log-file $O/launches_s1_4k_h2.csv python tools/profile_step.py --size 4096 --matrix s1 > $O/ncu_launch_h2.log 2>&1
timeout 300 python tools/select_bench.py > $O/select_bench_h2.txt 2>&1; tail -4 $O/select_bench_h2.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_h2.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["check"]["parity_all_configs"], d["roofline"]["achieved"], d["launches_per_step"], {k: round(v, 3) for k, v in d["phase_ms"].items()})
for k, v in d["other_configs"].items(): print("  ", k, v["value"], v["launches_per_step"], v["check"]["parity"], {a: round(b, 3) for a, b in v["phase_ms"].items()})
print("  eig-only", d["eigenvalues_only"]["value"], "select", d["selected_mode"]["device_s_per_solve"], "e2e", d["e2e"]["value"])
PY
