#!/bin/bash
# r02 call I (1 GPU): final-code validation: suite, smoke, dense bench, bench (driver arguments), reference arm, ncu traffic of the
# top-merge GEMM with the wider super-columns.
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_i.txt 2>&1; echo "pytest rc $?" >> $O/pytest_i.txt; tail -3 $O/pytest_i.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_i.txt 2>&1; echo "smoke rc $?" >> $O/smoke_i.txt; tail -2 $O/smoke_i.txt
timeout 600 python tools/dense_bench.py 4096 8192 16384 > $O/dense_bench_i.txt 2>&1; tail -3 $O/dense_bench_i.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_i.json 2> $O/bench_i.err; echo "bench rc $?" >> $O/bench_i.err; tail -1 $O/bench_i.err
python tools/profile_step.py --size 16384 --matrix goe > $O/prof_plain_i.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 5 -c 1 -o $O/prof_gemm_i python tools/profile_step.py --size 16384 --matrix goe > $O/ncu_full_i.log 2>&1
ncu -i $O/prof_gemm_i.ncu-rep --page raw --csv > $O/prof_gemm_i_raw.csv 2>/dev/null; rm -f $O/prof_gemm_i.ncu-rep
python - <<'PY'
import json, csv
d=json.loads(open("gpurun_out/r02/bench_i.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["check"]["parity_all_configs"], d["roofline"]["achieved"], d["roofline"]["frac"], d["launches_per_step"], {k: round(v, 3) for k, v in d["phase_ms"].items()})
for k, v in d["other_configs"].items(): print("  ", k, v["value"], v["launches_per_step"], v["check"]["parity"], v["roofline"]["achieved"])
print("  eig-only", d["eigenvalues_only"], "select", d["selected_mode"]["device_s_per_solve"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
rows=list(csv.reader(open("gpurun_out/r02/prof_gemm_i_raw.csv"))); h=rows[0]
for r in rows[2:]: print("ncu gemm:", r[h.index("gpu__time_duration.sum")], "ms  dram read", r[h.index("dram__bytes_read.sum")], "GB  write", r[h.index("dram__bytes_write.sum")], "GB  dmma", r[h.index("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active")], " L2 hit", r[h.index("lts__t_sector_hit_rate.pct")])
PY
