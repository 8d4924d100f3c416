timeout 300 python tools/resid_variants.py > gpurun_out/resid_variants_h.jsonl 2> gpurun_out/resid_variants_h.err
python - <<'PY'
import json
for l in open("gpurun_out/resid_variants_h.jsonl"):
    j=json.loads(l); print(j["n"], {k:v["GBps"] for k,v in j.items() if k!="n"})
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r01h.txt; cat gpurun_out/pytest_r01h.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_h.json 2> gpurun_out/bench_default_h.err; tail -c 300 gpurun_out/bench_default_h.err
for m in wilk goe; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_h.json 2> gpurun_out/bench_${m}16k_h.err; done
timeout 300 python tools/profile_step.py --size 65536 --matrix goe --select 16 --reps 0 > gpurun_out/prof_plain_sel_h.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cauchy_apply|secular_kernel|loewner_tiled|norms_tiled|rowgemv_tiled|compact_scan|RankLive' -c 80 -o gpurun_out/prof_select_goe65536 python tools/profile_step.py --size 65536 --matrix goe --select 16 --reps 0 > gpurun_out/ncu_sel_h.log 2>&1
tail -2 gpurun_out/ncu_sel_h.log
