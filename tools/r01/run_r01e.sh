timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r01e.txt; cat gpurun_out/pytest_r01e.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_e.json 2> gpurun_out/bench_default_e.err; tail -c 300 gpurun_out/bench_default_e.err
for m in wilk goe; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_e.json 2> gpurun_out/bench_${m}16k_e.err; done
timeout 300 python tools/select_bench.py --sizes 16384,65536 --ks 1,16,64 > gpurun_out/select_bench_e.jsonl 2> gpurun_out/select_bench_e.err
timeout 300 python tools/profile_step.py --size 65536 --matrix goe --select 16 > gpurun_out/prof_plain_sel_e.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_select_goe_65536_e.csv python tools/profile_step.py --size 65536 --matrix goe --select 16 --reps 0 > gpurun_out/ncu_sel_e.log 2>&1
cat gpurun_out/prof_plain_sel_e.log; cat gpurun_out/select_bench_e.jsonl | cut -c1-200
