timeout 300 python tools/resid_variants.py > gpurun_out/resid_variants.jsonl 2> gpurun_out/resid_variants.err; cat gpurun_out/resid_variants.jsonl | cut -c1-700
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r01f.txt; cat gpurun_out/pytest_r01f.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_f.json 2> gpurun_out/bench_default_f.err; tail -c 300 gpurun_out/bench_default_f.err
for m in wilk goe; do timeout 300 python bench.py --steps 3 --warmup 3 --matrix $m --size 16384 --no-cpu-baseline > gpurun_out/bench_${m}16k_f.json 2> gpurun_out/bench_${m}16k_f.err; done
timeout 300 python tools/select_bench.py --sizes 4096,16384,65536 --ks 1,16,64 > gpurun_out/select_bench_f.jsonl 2> gpurun_out/select_bench_f.err
cat gpurun_out/select_bench_f.jsonl | cut -c1-200
