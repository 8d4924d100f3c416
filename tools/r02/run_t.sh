#!/bin/bash
# r02 call T (1 GPU): validation of the final code: suite, smoke, dense bench, bench (driver arguments), launch lists of GOE n=16384 and
# `-s 1 -n 4096`, ncu --set full of the top-merge GEMM (DRAM traffic of the roofline entry) and of the vector kernels of the last solve.
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_t.txt 2>&1; echo "pytest rc $?" >> $O/pytest_t.txt; tail -3 $O/pytest_t.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_t.txt 2>&1; echo "smoke rc $?" >> $O/smoke_t.txt; tail -2 $O/smoke_t.txt
timeout 600 python tools/dense_bench.py 4096 8192 > $O/dense_bench_t.txt 2>&1; tail -3 $O/dense_bench_t.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_t.json 2> $O/bench_t.err; echo "bench rc $?" >> $O/bench_t.err; tail -1 $O/bench_t.err
timeout 300 python tools/select_bench.py --sizes 16384,65536 --ks 1,16 > $O/select_bench_t.jsonl 2> $O/select_bench_t.err
python tools/profile_step.py --size 16384 --matrix goe > $O/prof_plain_t.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_goe16k_t.csv python tools/profile_step.py --size 16384 --matrix goe > $O/ncu_launch_t.log 2>&1
python tools/profile_step.py --size 4096 --matrix s1 > $O/prof_plain_s1_t.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_s1_4k_t.csv python tools/profile_step.py --size 4096 --matrix s1 > $O/ncu_launch_s1_t.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 5 -c 1 -o $O/prof_gemm_t python tools/profile_step.py --size 16384 --matrix goe > $O/ncu_full_t.log 2>&1
ncu -i $O/prof_gemm_t.ncu-rep --page raw --csv > $O/prof_gemm_t_raw.csv 2>/dev/null; rm -f $O/prof_gemm_t.ncu-rep
ncu --set full --clock-control none -k regex:"ugen|secular|loewner_tiled|norms_tiled|pack_kernel|residual|rank_tiled|compact_scan" -s 48 -c 48 -o $O/prof_vec_t python tools/profile_step.py --size 16384 --matrix goe > $O/ncu_full_vec_t.log 2>&1
ncu -i $O/prof_vec_t.ncu-rep --page raw --csv > $O/prof_vec_t_raw.csv 2>/dev/null; rm -f $O/prof_vec_t.ncu-rep
python - <<'PY'
import json, csv
d=json.loads(open("gpurun_out/r02/bench_t.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["check"]["parity_all_configs"], d["roofline"]["achieved"], d["roofline"]["frac"], d["launches_per_step"], {k: round(v, 3) for k, v in d["phase_ms"].items()})
for k, v in d["other_configs"].items(): print("  ", k, v["value"], v["launches_per_step"], v["check"]["parity"], v["roofline"]["achieved"])
print("  eig-only", d["eigenvalues_only"], "select", d["selected_mode"]["device_s_per_solve"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
rows=list(csv.reader(open("gpurun_out/r02/prof_gemm_t_raw.csv"))); h=rows[0]
for r in rows[2:]: print("ncu gemm:", r[h.index("gpu__time_duration.sum")], "ms  dram read", r[h.index("dram__bytes_read.sum")], "GB  write", r[h.index("dram__bytes_write.sum")], "GB  dmma", r[h.index("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active")], " L2 hit", r[h.index("lts__t_sector_hit_rate.pct")])
rows=list(csv.reader(open("gpurun_out/r02/prof_vec_t_raw.csv"))); h=rows[0]
for r in rows[2:]: print("ncu vec:", r[h.index("Kernel Name")][:24], r[h.index("gpu__time_duration.sum")], "fp64", r[h.index("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")], "issue", r[h.index("smsp__issue_active.avg.pct")], "dram r/w", r[h.index("dram__bytes_read.sum")], r[h.index("dram__bytes_write.sum")])
PY
