out=gpurun_out/quad_sweep.log; : > $out
run() { echo "## R=$1 W=$2" >> $out; SGBPO_ROLLOUT_R=$1 SGBPO_ROLLOUT_WARPS=$2 timeout 200 python bench.py --no-cpu-baseline --steps 8 --warmup 3 --env BalanceQuadrotor --safeguard boundary_projection --num-envs 4096 --horizon 43 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])" >> $out 2>&1; }
run 2 14
run 2 4
run 4 4
run 8 4
run 8 2
cat $out
