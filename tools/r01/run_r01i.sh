timeout 300 python tools/diag_select.py 2>&1 | tail -12
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_r01i.txt; cat gpurun_out/pytest_r01i.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_i.json 2> gpurun_out/bench_default_i.err; tail -c 300 gpurun_out/bench_default_i.err
for m in wilk goe; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_i.json 2> gpurun_out/bench_${m}16k_i.err; done
timeout 300 python tools/select_bench.py --sizes 4096,16384,65536 --ks 1,16,64 > gpurun_out/select_bench_i.jsonl 2> gpurun_out/select_bench_i.err
du -sh gpurun_out
