#!/bin/bash
# round 2, GPU call 14 (8 GPUs): the scaling bench exactly as the driver launches it, with per-phase timing of rank 0
bash scripts/multi_gpu.sh 8 n8
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n8.json"))
print("phases", d.get("phases_rank0"))
print("clocks", d.get("clocks"))
PY
