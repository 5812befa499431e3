#!/bin/bash
# round 2, GPU call 23 (8 GPUs): the scaling bench exactly as the driver launches it, final code
bash scripts/multi_gpu.sh 8 n8_final
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n8_final.json"))
print("phases", d.get("phases_rank0"))
print("clocks", d.get("clocks"))
PY
