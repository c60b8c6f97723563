import os, sys
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import torch, numpy as np
from rcs_b200.envs.creators import FR3SimplePickUpSimEnvCreator
N=4096
env = FR3SimplePickUpSimEnvCreator()(num_envs=N)
obs,_=env.reset()
b=env.sim.batch
gen=torch.Generator(device=b.dev).manual_seed(1)
scale=torch.tensor([0.01]*3+[0.05]*3,dtype=torch.float64,device=b.dev)
from collections import Counter
for i in range(12):
    if i%10==0: env.reset()
    a={"xyzrpy":(torch.rand((N,6),dtype=torch.float64,device=b.dev,generator=gen)*2-1)*scale,"gripper":torch.randint(0,2,(N,),device=b.dev,generator=gen).to(torch.float64)}
    env.step(a)
    si=b.si.cpu().numpy()
    print(i, "solver_iter", sorted(Counter(si[:,16]).items())[:8], "ncon", sorted(Counter(si[:,14]).items()), "nefc", sorted(Counter(si[:,15]).items())[:6])
