import numpy as np, torch, sys
sys.path.insert(0,'.')
from tests.test_gpu_burger import make_env
g=np.load('tests/golden/burger_steps.npz')
V = g["direct/v"]
for n in (1,5):
    env, M = make_env("direct", g, B=2)
    env.IC(v0=V[[0, 0]])
    acts = np.zeros((2, 32)); acts[1] = 1e200
    st, _ = env.step_n(acts, n)
    print(n, env.status.cpu().numpy(), env.ioutnum_all.cpu().numpy(), st[1][:4].cpu().numpy(), env.v[1][:3].cpu().numpy())
