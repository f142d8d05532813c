#!/bin/bash
# Round-2 bench lines of the other BASELINE.json configs and of the size sweep, one JSON line each -> gpurun_out/r2_configs.log
out=gpurun_out/r2_configs.log; : > $out
run() { echo "## $*" >> $out; timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 3 "$@" 2>/dev/null | tail -1 >> $out; }
run --env BalancePendulum --safeguard ray_mask --num-envs 64 --horizon 64
run --env SwingUpCartPole --safeguard boundary_projection --num-envs 2048 --horizon 64
run --env SwingUpCartPole --safeguard ray_mask --num-envs 2048 --horizon 64
run --env BalanceCartPole --safeguard ray_mask --num-envs 4096 --horizon 64
for n in 4096 8192 16384 32768 65536; do run --num-envs $n; done
for e in 1 2 3; do for n in 4096 8192 16384; do echo "## engine $e" >> $out; SGBPO_ROLLOUT_ENGINE=$e run --num-envs $n; done; done
run --env ManageHousehold --safeguard boundary_projection --num-envs 8192 --horizon 550 --steps 5
run --env NavigateSeeker --safeguard boundary_projection --num-envs 65536 --horizon 256 --steps 5
run --env BalanceQuadrotor --safeguard boundary_projection --num-envs 1024 --horizon 16 --steps 2
