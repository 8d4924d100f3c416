#!/bin/bash
# r02 call J (1 GPU): suite (incl. the reciprocal self-test), bench without the CPU leg, eigenvalue phase at n=65536,
# GEMM L2-hint / super-column sweep: times plain, DRAM traffic under ncu (4 metrics, every dgemm_tma launch).
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_j.txt 2>&1; echo "pytest rc $?" >> $O/pytest_j.txt; tail -4 $O/pytest_j.txt
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_j.json 2> $O/bench_j.err; echo "bench rc $?" >> $O/bench_j.err; tail -1 $O/bench_j.err
timeout 300 python tools/select_bench.py --sizes 65536 --ks 1,16 > $O/select_bench_j.jsonl 2> $O/select_bench_j.err; cat $O/select_bench_j.jsonl | cut -c1-220
timeout 300 python tools/gemm_hint_sweep.py --reps 3 > $O/hint_sweep_j.jsonl 2> $O/hint_sweep_j.err; cat $O/hint_sweep_j.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --print-units base \
  -k regex:dgemm_tma --csv --log-file $O/hint_sweep_ncu_j.csv python tools/gemm_hint_sweep.py --reps 1 > $O/hint_sweep_ncu_j.log 2>&1; tail -2 $O/hint_sweep_ncu_j.log
python - <<'PY'
import json, csv
d=json.loads(open("gpurun_out/r02/bench_j.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["check"]["parity_all_configs"], d["roofline"]["achieved"], d["roofline"]["frac"], d["launches_per_step"], {k: round(v, 3) for k, v in d["phase_ms"].items()})
for k, v in d["other_configs"].items(): print("  ", k, v["value"], v["launches_per_step"], v["check"]["parity"], {a: round(b, 3) for a, b in v["phase_ms"].items()})
print("  eig-only", d["eigenvalues_only"], "select", d["selected_mode"]["device_s_per_solve"], "e2e", d["e2e"]["value"])
rows=list(csv.reader(open("gpurun_out/r02/hint_sweep_ncu_j.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]; h=rows[hi]
idc, mn, mv = h.index("ID"), h.index("Metric Name"), h.index("Metric Value")
L={}
for r in rows[hi+1:]:
    if len(r) > mv: L.setdefault(int(r[idc]), {})[r[mn]] = float(r[mv].replace(",", ""))
ids=sorted(L); per=12
for ci in range(len(ids)//per):
    top=max((L[i] for i in ids[ci*per:(ci+1)*per]), key=lambda x: x["gpu__time_duration.sum"])
    print("config", ci, {k: round(v, 3) for k, v in top.items()})
PY
