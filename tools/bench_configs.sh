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
# configs[4]: wide policies, not covered by the fused engine yet -> eager per-step kernels + torch MLP (slow: ~1 min each)
if [ "$1" = "wide" ]; then
  run2() { echo "## $*" >> $out; timeout 500 python bench.py --no-cpu-baseline --steps 3 --warmup 3 "$@" 2>&1 | tail -1 >> $out; }
  run2 --env ManageHousehold --safeguard boundary_projection --num-envs 4096 --horizon 550 --policy-arch 512,512,512 --vf-arch 512,512,512
  run2 --env NavigateSeeker --safeguard boundary_projection --num-envs 4096 --horizon 256 --policy-arch 436,436,436 --vf-arch 436,436,436
fi
