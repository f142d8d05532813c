import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from test_gpu_fused_rollout import build, Recorder, Replay
from safegbpo_b200 import rng
torch.set_default_dtype(torch.float64); torch.set_default_device("cuda:0"); dev=torch.device("cuda:0")
N,H=40,12
torch.manual_seed(3)
env_e, wrap_e, algo_e = build("BalanceQuadrotor","none",{},32,6,N,H,"off",dev)
env_f, wrap_f, algo_f = build("BalanceQuadrotor","none",{},32,6,N,H,"on",dev)
algo_f.policy.load_state_dict(algo_e.policy.state_dict()); algo_f.target_value_function.load_state_dict(algo_e.target_value_function.state_dict())
half = env_e.state_set.generator[0].diagonal()
s0 = env_e.state_set.center[0] + (torch.rand(N, 6)*2-1)*half*0.6
for env, algo in ((env_e, algo_e),(env_f, algo_f)):
    env.state, env.steps, env._reset_pending = s0.clone(), 2, False
    env.goal = s0[:, :2].clone()
    o0 = env.observation
    algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
rec=Recorder(); rng.set_source(rec)
algo_e.buffer.reset(); algo_e.collect_trajectories()
print("draws", [(k,tuple(d.shape)) for k,d in rec.draws][:8], len(rec.draws))
rng.set_source(Replay(rec.draws, dev))
algo_f.buffer.reset(); algo_f.collect_trajectories()
oe, of = algo_e.buffer.observations.tensor.detach(), algo_f.buffer.observations.tensor
ae, af = algo_e.buffer.actions.tensor.detach(), algo_f.buffer.actions.tensor
for t in range(H+1):
    print(t, "obs diff per col", (oe[t]-of[t]).abs().max(0).values.cpu().numpy().round(6), "act", float((ae[t]-af[t]).abs().max()))
print("goal eager", env_e.goal[:2], "fused", env_f.goal[:2])
print("s0[0]", s0[0].cpu().numpy())
print("eager obs1[0]", oe[1,0].cpu().numpy())
print("fused obs1[0]", of[1,0].cpu().numpy())
print("action0[0]", ae[0,0].cpu().numpy(), af[0,0].cpu().numpy())
print("noise draw0[0]", rec.draws[1][1][0].cpu().numpy())
st = algo_f._fused_engine.last["states"]
print("fused states0[0]", st[0,0].cpu().numpy(), "states1[0]", st[1,0].cpu().numpy())
worst = (oe[1]-of[1]).abs().max(1).values.argmax().item()
print("worst env", worst, "eager", oe[1,worst].cpu().numpy(), "fused", of[1,worst].cpu().numpy(), "s0", s0[worst].cpu().numpy(), "a", ae[0,worst].cpu().numpy())
print("per-env x diff", (oe[1,:,0]-of[1,:,0]).abs().cpu().numpy().round(3))
