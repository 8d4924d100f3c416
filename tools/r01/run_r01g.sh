timeout 300 python tools/resid_variants.py > gpurun_out/resid_variants_g.jsonl 2> gpurun_out/resid_variants_g.err
CUPPEN_LDPAD=48 timeout 300 python tools/resid_variants.py > gpurun_out/resid_variants_g_pad48.jsonl 2>> gpurun_out/resid_variants_g.err
python - <<'PY'
import json
for f in ("gpurun_out/resid_variants_g.jsonl","gpurun_out/resid_variants_g_pad48.jsonl"):
    for l in open(f):
        j=json.loads(l); print(f[-14:], j["n"], {k:v["GBps"] for k,v in j.items() if k!="n"})
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r01g.txt; cat gpurun_out/pytest_r01g.txt
