#!/bin/bash
# r02 call H3 (1 GPU): U kernel with batched independent loads (dorgv), fused front end with CTA size by merge count; leaf size 8 vs 16.
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_h3.txt 2>&1; echo "pytest rc $?" >> $O/pytest_h3.txt; tail -4 $O/pytest_h3.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_h3.json 2> $O/bench_h3.err; echo "bench rc $?" >> $O/bench_h3.err; tail -1 $O/bench_h3.err
for L in 8 32; do CUPPEN_LEAF=$L timeout 300 python bench.py --workload s1_4k --steps 10 --warmup 3 --no-cpu-baseline --select 0 > $O/bench_h3_leaf$L.json 2> $O/bench_h3_leaf$L.err; done
python tools/profile_step.py --size 16384 --matrix goe > $O/prof_plain_h3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 86 -c 90 --csv --log-file $O/launches_goe16k_h3.csv python tools/profile_step.py --size 16384 --matrix goe > $O/ncu_launch_h3.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_h3.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["check"]["parity_all_configs"], d["roofline"]["achieved"], d["launches_per_step"], {k: round(v, 3) for k, v in d["phase_ms"].items()})
for k, v in d["other_configs"].items(): print("  ", k, v["value"], v["launches_per_step"], v["check"]["parity"], {a: round(b, 3) for a, b in v["phase_ms"].items()})
print("  eig-only", d["eigenvalues_only"]["value"], "select", d["selected_mode"]["device_s_per_solve"], "e2e", d["e2e"]["value"])
for L in (8, 32):
    try:
        x=json.loads(open("gpurun_out/r02/bench_h3_leaf%d.json" % L).read().strip().splitlines()[-1]); print("  leaf", L, x["value"], x["check"]["parity"], {k: round(v, 3) for k, v in x["phase_ms"].items()})
    except Exception as e: print("leaf", L, "failed", e)
PY
