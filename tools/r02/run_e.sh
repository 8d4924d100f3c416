#!/bin/bash
# r02 call E (1 GPU): suite + bench after the ugen / work-list / residual / tensor-map changes, launch list and a full ncu
# capture of the vector kernels on the headline workload.
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_e.txt 2>&1; echo "pytest rc $?" >> $O/pytest_e.txt; tail -3 $O/pytest_e.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_e.json 2> $O/bench_e.err; echo "bench rc $?" >> $O/bench_e.err; tail -1 $O/bench_e.err
python tools/profile_step.py --size 16384 --matrix goe > $O/prof_plain_e.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 103 -c 110 --csv --log-file $O/launches_goe16k_e.csv python tools/profile_step.py --size 16384 --matrix goe > $O/ncu_launch_e.log 2>&1
python tools/profile_step.py --size 16384 --matrix goe > $O/prof_plain_e2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ugen|secular|loewner_tiled|norms_tiled|pack_kernel|residual|rank_tiled|compact_scan" -s 52 -c 52 -o $O/prof_vec_goe16k python tools/profile_step.py --size 16384 --matrix goe > $O/ncu_full_e.log 2>&1
ncu -i $O/prof_vec_goe16k.ncu-rep --page raw --csv > $O/prof_vec_goe16k_raw.csv 2>/dev/null; rm -f $O/prof_vec_goe16k.ncu-rep
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_e.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["check"]["parity"], d["roofline"]["achieved"], d["roofline"]["frac"], {k: round(v, 3) for k, v in d["phase_ms"].items()})
for k, v in d["other_configs"].items(): print("  ", k, v["value"], v["launches_per_step"], v["check"]["parity"], {a: round(b, 3) for a, b in v["phase_ms"].items()})
PY
