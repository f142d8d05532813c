#!/bin/bash
# The other BASELINE.json configs (per-GPU shard sizes), one bench line each -> gpurun_out/configs.log
out=gpurun_out/configs.log; : > $out
run() { echo "## $*" >> $out; timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 3 "$@" 2>&1 | tail -1 >> $out; }
run --env BalancePendulum --safeguard ray_mask --num-envs 64 --horizon 64
run --env SwingUpCartPole --safeguard boundary_projection --num-envs 2048 --horizon 64
run --env SwingUpCartPole --safeguard ray_mask --num-envs 2048 --horizon 64
run --env BalanceQuadrotor --safeguard boundary_projection --num-envs 4096 --horizon 43
run --env BalanceCartPole --safeguard boundary_projection --num-envs 65536 --horizon 64
run --env BalanceCartPole --safeguard ray_mask --num-envs 4096 --horizon 64
