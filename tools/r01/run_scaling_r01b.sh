set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 4 8; do
  $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 --matrix goe --size 16384 > gpurun_out/scale2_goe_16384_g$N.json 2> gpurun_out/scale2_goe_16384_g$N.err
  tail -c 200 gpurun_out/scale2_goe_16384_g$N.json; tail -3 gpurun_out/scale2_goe_16384_g$N.err
done
$TR --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --steps 2 --warmup 3 --matrix goe --size 32768 > gpurun_out/scale2_goe_32768_g8.json 2> gpurun_out/scale2_goe_32768_g8.err
$TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 2 --warmup 3 --matrix goe --size 32768 --no-single-gpu-compare > gpurun_out/scale2_goe_32768_g4.json 2> gpurun_out/scale2_goe_32768_g4.err
$TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --steps 2 --warmup 3 --matrix goe --size 65536 > gpurun_out/scale2_goe_65536_g8.json 2> gpurun_out/scale2_goe_65536_g8.err
tail -c 300 gpurun_out/scale2_goe_65536_g8.json; tail -3 gpurun_out/scale2_goe_65536_g8.err
