timeout 60 ./cuppens -p 8 -g 1 -s 1 -n 2048 -e gpurun_out/u_g1.txt > gpurun_out/u_g1.log 2>&1
timeout 60 ./cuppens -p 8 -g 2 -s 1 -n 2048 -e gpurun_out/u_g2.txt > gpurun_out/u_g2.log 2>&1
timeout 60 ./cuppens -p 4 -g 2 -i tests/golden/tinyL.mtx -e gpurun_out/u_tiny.txt > gpurun_out/u_tiny.log 2>&1; tail -1 gpurun_out/u_tiny.log
python - <<'PY'
import numpy as np
a=np.loadtxt('gpurun_out/u_g1.txt'); b=np.loadtxt('gpurun_out/u_g2.txt')
print('lam diff', np.abs(a[:,0]-b[:,0]).max(), 'resid max', a[:,1].max(), b[:,1].max(), 'resid diff', np.abs(a[:,1]-b[:,1]).max())
PY
tail -2 gpurun_out/u_g2.log
