python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r01d.txt; cat gpurun_out/pytest_r01d.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default_d.json 2> gpurun_out/bench_default_d.err; tail -c 600 gpurun_out/bench_default_d.json
for m in wilk goe; do python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_d.json 2> gpurun_out/bench_${m}16k_d.err; done
python tools/profile_step.py --size 16384 --matrix wilk --orth > gpurun_out/prof_plain_wilk.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'residual_kernel|gram_check|pack_kernel|ugen_kernel' -s 20 -c 14 -o gpurun_out/prof_wilk16k_vec python tools/profile_step.py --size 16384 --matrix wilk --orth > gpurun_out/ncu_wilk.log 2>&1
python tools/profile_step.py --size 65536 --matrix goe --select 16 > gpurun_out/prof_plain_sel.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_select_goe_65536.csv python tools/profile_step.py --size 65536 --matrix goe --select 16 --reps 0 > gpurun_out/ncu_sel.log 2>&1
cat gpurun_out/prof_plain_wilk.log gpurun_out/prof_plain_sel.log
