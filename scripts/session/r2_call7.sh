#!/bin/bash
# round 2, GPU call 7: where does conv_tc5's per-tile time go?  (timing experiments: stores off / MMAs off)
mkdir -p gpurun_out
for d in 0 1 2 3; do
echo "== SAE_TC_DEBUG=$d"
SAE_TC_DEBUG=$d timeout 200 python scripts/conv_bench.py --only "s1" --dirs fprop 2>&1 | grep -v "^shape" | head -8
done
echo "== SAE_TC_DEBUG=0..3, TC5X=1"
for d in 0 1 2 3; do
SAE_TC5X=1 SAE_TC_DEBUG=$d timeout 200 python scripts/conv_bench.py --only "D/G" --dirs fprop 2>&1 | grep -v "^shape" | head -3
done
