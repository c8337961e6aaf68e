import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'oracle'))
import numpy as np, nmf_jl_b200 as NMF, nmf_oracle as O
for (p,n,k,iters) in [(1001,1030,200,3),(1024,896,64,4),(2048,1536,200,3)]:
    rng=np.random.default_rng(42)
    X=np.asfortranarray(rng.random((p,n)),dtype=np.float32)
    W0,H0=NMF.randinit(p,n,k,np.float32,normalize=True,rng=rng)
    W,H=W0.copy(order='F'),H0.copy(order='F')
    r=NMF.solve(NMF.GreedyCD(np.float32,maxiter=iters,tol=1e-9),X,W,H,engine='tc')
    Wo,Ho=W0.copy(order='F'),H0.copy(order='F')
    ro=O.solve(O.GreedyCD(np.float32,maxiter=iters,tol=1e-9),X,Wo,Ho)
    print(p,n,k,'single tc errObj',abs(float(r.objvalue)-float(ro.objvalue))/float(ro.objvalue), r.info['coordinate_updates'], ro.coordinate_updates, flush=True)
