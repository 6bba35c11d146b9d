"""Deviation of the opt-in fast kernels from the reference-exact oracle on BASELINE config 2's sweep (20 000 trajectories).
usage: python tools/fast_rhs_diff.py [compat]    8 = SDE_COMPAT_FAST_RHS (default), 16 = SDE_COMPAT_FAST_STAGES, 24 = both"""
import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, simplediffeq_b200 as sde, common as C, oracle_lib
oracle_lib.build()
n=20000
COMPAT=int(sys.argv[1]) if len(sys.argv)>1 else 8
print('compat',COMPAT)
u0,p=C.lorenz_sweep(n)
tg=sde.jl_range(0.0,1e-3,10.0)
o=oracle_lib.solve("lorenz","Tsit5",u0,p,0.0,10.0,1e-3,tgrid=tg,n_threads=16)
ref=np.ascontiguousarray(o.u[:,0,:])
g=sde.solve_arrays(sde.systems.lorenz, sde.GPUSimpleTsit5(), np.ascontiguousarray(u0.T), np.ascontiguousarray(p.T),(0.0,10.0),dt=1e-3,compat=COMPAT)
fu=np.ascontiguousarray(g["u"].T)
rel=np.abs(fu-ref)/np.maximum(np.abs(ref),1e-300)
mx=rel.max(axis=1)
print("max rel", rel.max(), "at", np.argmax(mx), "rho", p[np.argmax(mx)], ref[np.argmax(mx)], fu[np.argmax(mx)])
for q in (50,90,99,99.9,99.99): print(q, np.percentile(mx,q))
bad=np.flatnonzero(mx>1e-12); print(len(bad), "above 1e-12; rho range", p[bad,1].min() if len(bad) else None, p[bad,1].max() if len(bad) else None)
for band in (0.01,0.025,0.05):
    far=np.abs(p[:,1]-13.926)>band; print("max rel outside +-%g of rho=13.926:"%band, mx[far].max())
absd=np.abs(fu-ref).max(axis=1); print("max abs", absd.max(), "rho", p[np.argmax(absd),1])
