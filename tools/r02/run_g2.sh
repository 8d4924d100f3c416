#!/bin/bash
# r02 call G2 (1 GPU): the n=32768 / n=65536 extra keys of bench.py on one GPU (--big): validates the LAPACK subprocess,
# the golden check at 32768 and the accurate-rule run at 65536 before the 8-GPU run.
O=gpurun_out/r02; mkdir -p $O
timeout 1700 python bench.py --steps 3 --warmup 3 --big --no-cpu-baseline --select 0 > $O/bench_g2_big_1gpu.json 2> $O/bench_g2_big_1gpu.err; echo "bench rc $?" >> $O/bench_g2_big_1gpu.err; tail -3 $O/bench_g2_big_1gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_g2_big_1gpu.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["check"]["parity"], d["check"]["parity_all_configs"])
for k, v in d["other_configs"].items():
    print("  ", k, v["value"], v["check"]["parity"], {a: round(b, 3) for a, b in v["phase_ms"].items()}, v["roofline"]["achieved"], {a: b for a, b in v["check"].items() if "lambda" in a or "resid" in a or "orth" in a})
    if "accurate_rule" in v: print("     accurate rule:", v["accurate_rule"]["value"], v["accurate_rule"]["check"])
PY
