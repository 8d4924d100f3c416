#!/bin/bash
# strong-scaling runs on one 8xB200 box (profiles/README.md); usage: bash tools/run_scaling.sh [quick]
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
if [ "$1" == "quick" ]; then
  $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --matrix goe --size 4096 > gpurun_out/scale_goe_4096_g2.json 2> gpurun_out/scale_goe_4096_g2.err
  tail -c 600 gpurun_out/scale_goe_4096_g2.json; tail -5 gpurun_out/scale_goe_4096_g2.err
  $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/scale_goe_16384_g2.json 2> gpurun_out/scale_goe_16384_g2.err
  tail -c 600 gpurun_out/scale_goe_16384_g2.json; tail -5 gpurun_out/scale_goe_16384_g2.err
  ./cuppens -p 8 -g 2 -s 2 -n 2048 -e gpurun_out/cli_g2.txt | tail -4; ./cuppens -p 8 -s 2 -n 2048 -e gpurun_out/cli_g1.txt | tail -2
  python - <<'PY'
import numpy as np
a=np.loadtxt("gpurun_out/cli_g2.txt"); b=np.loadtxt("gpurun_out/cli_g1.txt")
print("cli -g 2 vs -g 1: max dlam %.2e max dresid %.2e" % (np.abs(a[:,0]-b[:,0]).max(), np.abs(a[:,1]-b[:,1]).max()))
PY
  exit 0
fi
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 --matrix goe --size 16384 > gpurun_out/scale_goe_16384_g$N.json 2> gpurun_out/scale_goe_16384_g$N.err
  tail -c 300 gpurun_out/scale_goe_16384_g$N.json; tail -3 gpurun_out/scale_goe_16384_g$N.err
done
for N in 4 8; do
  $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --steps 2 --warmup 3 --matrix goe --size 32768 > gpurun_out/scale_goe_32768_g$N.json 2> gpurun_out/scale_goe_32768_g$N.err
done
$TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --steps 2 --warmup 3 --matrix goe --size 65536 > gpurun_out/scale_goe_65536_g8.json 2> gpurun_out/scale_goe_65536_g8.err
$TR --nproc-per-node 8 --master-port 29539 bench.py --gpus 8 --steps 3 --warmup 3 --matrix s1 --size 4096 > gpurun_out/scale_s1_4096_g8.json 2> gpurun_out/scale_s1_4096_g8.err
tail -c 300 gpurun_out/scale_goe_65536_g8.json; tail -3 gpurun_out/scale_goe_65536_g8.err
