import numpy as np, torch, sys, os
sys.path.insert(0,'.')
from tests.test_gpu_burger import make_env
g=np.load('tests/golden/burger_steps.npz')
case = sys.argv[1] if len(sys.argv)>1 else "eddy_forced"
V, A = g[case+"/v"], g[case+"/actions"]
for B in (1,2,3,4,5):
    eb,M=make_env(case, g, B=B)
    for n in (2,3):
        eb.IC(v0=V[[0]*B]); eb.step_n(A[[0]*B] if M else None, n)
        v=eb.v.reshape(B,-1).cpu().numpy()
        d=[[float(np.abs(v[e]-V[j]).max()) for j in range(0,5)] for e in range(B)]
        print(case,'B',B,'n',n, ['%d:'%e+','.join('%.1e'%x for x in d[e]) for e in range(B)], eb.ioutnum_all.cpu().numpy())
