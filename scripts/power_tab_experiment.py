"""Oracle-side experiment behind jc_power_tab_kernel (csrc/jc_power.cu): the transfer function T(k) tabulated per
cosmology on a uniform ln k grid (spacing h), node slopes by 4th-order central differences, cubic Hermite evaluation --
error on C_ell and on single V entries against the exact formula, bench tracer set (10+10 bins, 100 ell, halofit).
Test infrastructure: runs the NumPy oracle only.  Output: profiles/r02_power_tab.md."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle import cl_oracle as o, scenarios as sc
scn = sc.scenario("cfg5", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(10,1.0), sc.lenses(10,1.0)])
prob = sc.flatten_spec(scn); ell=np.array(scn["ell"])
rows = sc.config5_cosmologies(65536)[:6]
exact = o.eisenstein_hu
def make_tab(h, mode="T"):
    def eh(c,k):
        k=np.asarray(k); 
        if k.size < 1000: return exact(c,k)   # halofit/romberg grids stay exact
        lk=np.log(k); lo=lk.min()-3*h; hi=lk.max()+3*h
        n=int(np.ceil((hi-lo)/h))+1
        x=lo+h*np.arange(n); f=exact(c,np.exp(x))
        if mode=="T2": f=f*f
        d=np.zeros(n); d[2:-2]=(-f[4:]+8*f[3:-1]-8*f[1:-3]+f[:-4])/12.0  # h*f'
        u=(lk-lo)/h; i=np.floor(u).astype(int); i=np.clip(i,2,n-4); t=u-i
        f0,f1,d0,d1=f[i],f[i+1],d[i],d[i+1]
        D=f1-f0; a=d0+d1-2*D; b=D-d0-a
        r=f0+t*(d0+t*(b+t*a))
        return np.sqrt(r) if mode=="T2" else r
    return eh
ref=[o.angular_cl(r,ell,prob) for r in rows]
for mode in ("T",):
  for h in (0.001,0.002,0.0025,0.003,0.004,0.006):
    o.eisenstein_hu=make_tab(h,mode)
    errs=[np.max(np.abs(o.angular_cl(r,ell,prob)/ref[i]-1)) for i,r in enumerate(rows)]
    print(mode,h,"max rel err C_l", max(errs), "nodes", int(14.5/h))
o.eisenstein_hu=exact
print("pointwise V")
for h in (0.00235,0.003,0.004):
    st0={}; o.eisenstein_hu=exact; o.angular_cl(rows[0],ell,prob,stages=st0)
    st1={}; o.eisenstein_hu=make_tab(h); o.angular_cl(rows[0],ell,prob,stages=st1)
    rel=np.abs(st1["V"]/st0["V"]-1)
    print(h, "max rel V", rel.max(), "at (l,n)", np.unravel_index(rel.argmax(), rel.shape), "99.9pct", np.quantile(rel,0.999))
o.eisenstein_hu=exact
