#!/bin/bash
# usage: tools/quick_bench.sh <tag> [extra bench args] - build tests + a short bench, per-kernel table
tag=$1; shift
python -m pytest tests/test_build_gpu.py -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-candidates --e2e-steps 0 "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
tail -3 gpurun_out/$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/$tag.json"))
print(d["ms_per_step"], d["roofline"]["phase_ms"])
for k,v in d["roofline"]["kernels"].items(): print("  ", k, v["ms"])
PY
