set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_r01j.txt; cat gpurun_out/pytest_r01j.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default_j.json 2> gpurun_out/bench_default_j.err; tail -c 200 gpurun_out/bench_default_j.err
for m in wilk goe randu s2; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_j.json 2> gpurun_out/bench_${m}16k_j.err; done
timeout 300 python tools/select_bench.py --sizes 4096,16384,32768,65536 --ks 1,16,64 > gpurun_out/select_bench_j.jsonl 2> gpurun_out/select_bench_j.err
timeout 300 python tools/select_bench.py --sizes 4096 --matrix s1 --ks 1,16 >> gpurun_out/select_bench_j.jsonl 2>> gpurun_out/select_bench_j.err
timeout 120 python -c "
import json, symmetric_eigenvalue_b200.api as api
print(json.dumps(api.measure_fp64_mix()))" > gpurun_out/fp64_mix.json 2> gpurun_out/fp64_mix.err; cat gpurun_out/fp64_mix.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_s1_4096_j.csv python tools/profile_step.py --size 4096 --matrix s1 > gpurun_out/ncu_s1_j.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_select_goe_65536_j.csv python tools/profile_step.py --size 65536 --matrix goe --select 16 --reps 0 > gpurun_out/ncu_sel_j.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'cauchy_apply' -c 3 -o gpurun_out/prof_cauchy_goe16k python tools/profile_step.py --size 16384 --matrix goe --select 16 --reps 0 > gpurun_out/ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'rank_tiled|compact_scan|loewner_tiled|norms_tiled|residual_kernel|gram_check' -s 24 -c 6 -o gpurun_out/prof_vec_goe16k python tools/profile_step.py --size 16384 --matrix goe --reps 0 --orth > gpurun_out/ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'secular_kernel|rowgemv_tiled' -s 14 -c 2 -o gpurun_out/prof_secular_goe32k python tools/profile_step.py --size 32768 --matrix goe --select 8 --reps 0 > gpurun_out/ncu_c.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'dgemm' -c 16 -o gpurun_out/prof_dgemm_s1_4096 python tools/profile_step.py --size 4096 --matrix s1 --reps 0 > gpurun_out/ncu_d.log 2>&1
du -sh gpurun_out; ls -la gpurun_out
