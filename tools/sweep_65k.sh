#!/bin/bash
# Launch-shape sweep of the rollout kernels at 65 536 environments (env overrides, no rebuild) -> gpurun_out/sweep_65k.log
out=gpurun_out/sweep_65k.log; : > $out
run() { echo "## R=$1 W=$2" >> $out; SGBPO_ROLLOUT_R=$1 SGBPO_ROLLOUT_WARPS=$2 timeout 200 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --num-envs 65536 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value']/1e6,1), d['roofline']['kernel_ms'])" >> $out 2>&1; }
echo "## default" >> $out; timeout 200 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --num-envs 65536 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value']/1e6,1), d['roofline']['kernel_ms'])" >> $out 2>&1
run 8 8
run 8 6
run 8 4
run 4 12
run 4 8
cat $out
