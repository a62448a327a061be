import sys, time, numpy as np
sys.path.insert(0, ".")
import jax_cosmo_b200 as jc
from oracle import scenarios as sc, cl_oracle as o
src=[sc.smail(1.0,2.0,0.3+0.07*i,1.5) for i in range(9)]
lns=[sc.smail(2.0,4.0,0.25+0.06*i,2.0) for i in range(8)]
spec=[sc.wl(src), sc.nc(lns,[sc.bias("constant",1.0+0.05*i) for i in range(8)])]
for L in (1500, 2600):
    ell=np.logspace(0.5, 4, L)
    scn=sc.scenario("x", sc.WCDM, ell, spec)
    probes=sc.build_probes(scn, jc); cosmo=sc.build_cosmo(scn, jc)
    t=time.time(); cl=jc.cl.angular_cl(cosmo, ell, probes); t1=time.time()-t
    ref=o.angular_cl(sc.cosmo_row(scn["cosmo"]), ell, sc.flatten_spec(scn))
    print(L, cl.shape, "max rel err %.2e"%np.max(np.abs(cl-ref)/np.abs(ref)), "gpu call %.1f ms"%(t1*1e3))
