TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_default_g2_n.json 2> gpurun_out/bench_default_g2_n.err
tail -c 1500 gpurun_out/bench_default_g2_n.json; tail -3 gpurun_out/bench_default_g2_n.err
timeout 300 $TR --nproc-per-node 2 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-budget 12 > gpurun_out/bench_ref_g2_n.json 2> gpurun_out/bench_ref_g2_n.err; tail -c 600 gpurun_out/bench_ref_g2_n.json
