import sys; sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
import particlerobotsimulations_b200 as prs
from oracle import binding as ob
from tests import util
from tests.util import Dev
name, steps = sys.argv[1], int(sys.argv[2])
randv = len(sys.argv) > 3
p,o = util.cfg(name); s = util.oracle_state_after(p,o,steps)
n=p.nCells; dt=o.timestep
spos,svel,srad = s.get("sortedPos"), s.get("sortedVel"), s.get("sortedRad")
idx,cs,ce = s.get("index"), s.get("cellStart"), s.get("cellEnd"); fr0=s.get("absForce_r")
if randv:
    rng=np.random.default_rng(1); svel=(svel+rng.standard_normal(svel.shape).astype(np.float32)*0.05).astype(np.float32); srad=(srad+rng.random(n).astype(np.float32)*0.02).astype(np.float32)
def run(L):
    L.setParameters(C.byref(p))
    d=[Dev(np.zeros((n,2),np.float32)),Dev(np.zeros(n,np.float32)),Dev(fr0),Dev(spos),Dev(svel),Dev(srad),Dev(idx),Dev(cs),Dev(ce)]
    L.collide(*[x.ptr for x in d], n, p.numCells, dt)
    return d[0].get(), d[1].get(), d[2].get()
v,fa,fr = run(prs.lib()); vr,far,frr = run(util.refcuda())
def ulps(a,b): return np.abs(a.view(np.int32).astype(np.int64)-b.view(np.int32).astype(np.int64))
print("vel differing entries", int((v.view(np.uint32)!=vr.view(np.uint32)).sum()), "of", v.size, "max ulp", int(ulps(v,vr).max()))
print("fa differing", int((fa.view(np.uint32)!=far.view(np.uint32)).sum()), "max ulp", int(ulps(fa,far).max()))
print("fr differing", int((fr.view(np.uint32)!=frr.view(np.uint32)).sum()), "max ulp", int(ulps(fr,frr).max()))
bad = np.nonzero((v.view(np.uint32)!=vr.view(np.uint32)).any(1))[0][:8]
inv = np.empty(n,np.int64); inv[idx]=np.arange(n)
for b in bad:
    k=inv[b]; d=np.linalg.norm(spos-spos[k],axis=1); cd=srad+srad[k]; gap=d-cd
    m=(np.arange(n)!=k)&(d<0.8)
    print("robot",b,"v",v[b],vr[b],"fr",fr[b],frr[b],"fa",fa[b],far[b],"contacts",int((gap[m]<0).sum()),"near",int(((gap[m]>=0)&(gap[m]<0.0009)).sum()),"mid",int(((gap[m]>=0.0009)&(gap[m]<0.0019)).sum()),"far",int((gap[m]>=0.0019).sum()), "|v_in|", np.linalg.norm(svel[k]))
