import sys; sys.path.insert(0,".")
import numpy as np
from onesolver_b200 import Problem, capi, exhaustive
from oracle import binding as ob
for n in (1,2,3,31,32,33):
    rng=np.random.default_rng(n); a=rng.integers(-5,6,size=(n,n)).astype(float); q=np.triu(a,1); q=q+q.T+np.diag(np.diag(a))
    for mode in (0,1):
        for prec,dt in ((capi.SWEEP_F64,np.float64),(capi.SWEEP_F32,np.float32)):
            with Problem.dense(q, sweep_precision=prec) as p:
                r=p.anneal(np.linspace(0.5,5,6),6,37,mode=mode,want_energies=True,want_states=True)
            br,b,_,_=ob.replay_dense(q,np.linspace(0.5,5,6),6,37,mode=mode,dtype=dt)
            assert (b==r.best_states_packed).all(), (n,mode,dt)
            assert (ob.energy_packed(q,b)==r.best_energies).all()
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
        r=p.parallel_tempering([0.3,1.0,3.0],4,5,2,want_energies=True)
    if n<=20:
        st,e=exhaustive(q); assert r.energy>=e-1e-9
    print("n",n,"ok", r.energy)
