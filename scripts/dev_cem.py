import sys; sys.path.insert(0,'.')
import numpy as np
from oracle import cases
from tests.test_gpu_cem_std import _setup
from tests.util import stack_noise
case=cases.CEM_STD_CASES["cemstd_cheetah"]
p,orc,model,c=_setup(case)
np.random.seed(case["seed"]); obs=np.asarray(case["start_obs"],np.float64).copy()
orc.beginning_of_rollout(); p.begin_rollout()
n=c["num_simulated_trajectories"]
for s in range(2):
    m0,s0=orc.mean.copy(),orc.std.copy()
    tr=orc.get_action(obs)
    for i,it in enumerate(tr.iterations):
        u,_=stack_noise(it.noise); p.inject_noise(i,u,None)
    act=p.plan(obs)
    for i,it in enumerate(tr.iterations):
        a_dev=p.actions(i,n); d=np.abs(a_dev-it.actions)
        idx=np.unravel_index(d.argmax(),d.shape)
        u,_=stack_noise(it.noise)
        mean=m0 if i==0 else tr.iterations[i-1].mean; std=s0 if i==0 else tr.iterations[i-1].std
        lo=(orc.cfg.action_low-mean)/(std+1e-8); hi=(orc.cfg.action_high-mean)/(std+1e-8)
        t,dm=idx[1],idx[2]
        print(s,i,"max err",d.max(),"at",idx,"u",u[idx],"a",lo[t,dm],"b",hi[t,dm],"std",std[t,dm],"mean",mean[t,dm],"ref",it.actions[idx],"dev",a_dev[idx], "n>1e-5:",(d>1e-5).sum())
    obs=model.step(obs[None],tr.action[None])[0]
