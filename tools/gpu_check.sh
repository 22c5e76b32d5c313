#!/bin/bash
# One GPU-box pass: parity tests, K sweeps through the per-block API, the headline bench.  usage: gpu_check.sh <tag>
tag=${1:-x}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
T=gpurun_out/${tag}_trace.jsonl; : > $T
if [ -f open-headstage_b200/libohs_cuda_trace.so ]; then
  OHS_LIB_OVERRIDE=$PWD/open-headstage_b200/libohs_cuda_trace.so timeout 120 python tools/trace_k1.py 2 1 32 >> $T 2>gpurun_out/${tag}_trace.err
fi
for a in "2 1 64" "2 2 32" "2 4 16" "3 1 32" "5 1 16"; do timeout 120 python tools/trace_k1.py $a >> $T 2>>gpurun_out/${tag}_trace.err; done
cat $T; tail -3 gpurun_out/${tag}_trace.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "k1", d.get("per_block_api"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench.err").read()[-2000:])
PY
