import sys; sys.path.insert(0,'/root/repo')
from symmetric_eigenvalue_b200 import api
for (M,N,K) in [(8192,8192,8192),(4096,4096,4096),(8192,16384,8192),(2048,4096,2048)]:
    for v in (0,1):
        err, tf = api.selftest_gemm(v, M, N, K, reps=3)
        print("gemm variant", v, (M,N,K), "err %.2e"%err, "TF/s %.2f"%tf, flush=True)
