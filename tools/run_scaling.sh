set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 --matrix goe --size 16384 > gpurun_out/scale_goe_16384_g$N.json 2> gpurun_out/scale_goe_16384_g$N.err
  tail -c 400 gpurun_out/scale_goe_16384_g$N.json; tail -3 gpurun_out/scale_goe_16384_g$N.err
done
for N in 4 8; do
  $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --steps 2 --warmup 3 --matrix goe --size 32768 > gpurun_out/scale_goe_32768_g$N.json 2> gpurun_out/scale_goe_32768_g$N.err
done
$TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --steps 2 --warmup 3 --matrix goe --size 65536 > gpurun_out/scale_goe_65536_g8.json 2> gpurun_out/scale_goe_65536_g8.err
$TR --nproc-per-node 8 --master-port 29539 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/scale_s1_4096_g8.json 2> gpurun_out/scale_s1_4096_g8.err
tail -c 300 gpurun_out/scale_goe_65536_g8.json; tail -3 gpurun_out/scale_goe_65536_g8.err
