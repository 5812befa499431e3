#!/bin/bash
mkdir -p gpurun_out
for m in 2 1; do
echo "==== wgrad shared-window mode $m: conv tests"
SAE_WGRAD_WINDOW=$m timeout 300 python -m pytest tests -m gpu -q --timeout 120 -k "conv_fprop_dgrad_wgrad or adjointness" 2>&1 | grep -v "^E   *+\|^E  *where" | tail -7
done
