import sys, os, time
sys.path.insert(0, ".")
import numpy as np
import tci_b200 as T
ld=[10]*8
f=T.BuiltinTarget(T.LORENTZ,[1.0],ld)
rng=np.random.default_rng(0)
I=np.stack([rng.integers(1,11,120) for _ in range(4)],axis=1); J=np.stack([rng.integers(1,11,120) for _ in range(4)],axis=1)
for _ in range(5): f.bond_update(I,J,maxrank=12,abstol=1e-9)
ctx=f.ctx; ctx.timers(reset=True)
t0=time.perf_counter()
for _ in range(200): f.bond_update(I,J,maxrank=12,abstol=1e-9)
dt=(time.perf_counter()-t0)/200*1e6
tm=ctx.timers()
print(f"bond_update 120x120 r=12: {dt:.1f} us per call; stage timers per call: pi {tm['pi_eval']/200*1e3:.1f} us, rrlu stage {tm['rrlu']/200*1e3:.1f} us, rrlu kernel {tm['rrlu_kernel']/200*1e3:.1f} us", file=sys.stderr)
