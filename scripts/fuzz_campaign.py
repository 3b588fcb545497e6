#!/usr/bin/env python
"""Extended differential fuzz of the DEVICE ENGINE SOURCE (host build, tests/sim) against the oracle: the SURVEY A.8 domain with random
panel heights for both drivers (thread-per-pair, warp-cooperative) + long pairs with wide interior ranges (the interior fast step).
usage: fuzz_campaign.py <seed> <iterations>   (test infrastructure; the CPU suite runs a fixed, smaller version of this)"""
import sys, time; import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, harness as H, fuzzgen as F
from test_oracle import fuzz_batches
seed=int(sys.argv[1]); n_iter=int(sys.argv[2])
rng=np.random.default_rng(seed)
t0=time.time(); n=0; bad=0
def ring_ok(P,qs,ts):
    """the ring schedule serves pairs whose effective band is at most 512 (ksw2_pair.cuh KS_RING_MAX_W); exts2 has no band"""
    if P.kind==2: return all(max(len(q),len(t))<=512 for q,t in zip(qs,ts))
    return all((P.w if 0<=P.w<=max(len(q),len(t)) else max(len(q),len(t)))<=512 for q,t in zip(qs,ts))
def compare(P,qs,ts,js,panel,fs):
    global bad
    a=H.run_cpu("oracle",P,qs,ts,js); b=H.run_sim(P,qs,ts,js,panel=panel,force_smode=fs)
    ok=np.array_equal(a[0][:,:11],b[0][:,:11]) and all(np.array_equal(x,y) for x,y in zip(a[1],b[1]))
    if not ok:
        bad+=1; print("MISMATCH",P.kind,hex(P.flag),P.w,P.zdrop,P.q,P.e,P.q2,P.e2,panel,fs,[len(x) for x in qs],[len(x) for x in ts],flush=True)
# 1) standard fuzz domain, random seeds, both drivers
for P,qs,ts,js in fuzz_batches(seed,n_iter):
    compare(P,qs,ts,js,int(rng.choice([1,2,3,5,7,15,16,24,32,64,1000])),int(rng.integers(0,4)))
    if n%3==0: compare(P,qs,ts,js,-int(rng.choice([1,2,5,16,33,128])),int(rng.integers(0,4)))
    if n%5==0: compare(P,qs,ts,js,-(100000+int(rng.choice([1,7,40,200]))),int(rng.integers(0,4)))
    if n%2==0 and ring_ok(P,qs,ts): compare(P,qs,ts,js,-200000,int(rng.integers(0,4)))
    n+=1
# 2) longer pairs (wide interior ranges), varied bands
for it in range(n_iter//10):
    kind=["extz2","extd2","exts2"][it%3]
    tl=int(rng.integers(200,1500)); t=rng.integers(0,4,tl).astype(np.uint8)
    q=H.mutate(rng,t,sub=float(rng.uniform(0.01,0.15)),ins=float(rng.uniform(0,0.05)),dele=float(rng.uniform(0,0.05)))
    if rng.random()<0.3: q=q[:max(1,len(q)-int(rng.integers(1,200)))]
    if rng.random()<0.3: q=np.concatenate([q[:len(q)//2],rng.integers(0,4,len(q)//2).astype(np.uint8)])
    if rng.random()<0.3: t=t.copy(); t[int(rng.integers(0,tl)):][:5]=4
    a,b=F.AB[rng.integers(len(F.AB))]; mat=H.simple_mat(5,a,b,0 if rng.random()<0.7 else -1)
    w=int(rng.choice([-1,20,50,100,130,257,500])); zd=int(rng.choice([-1,50,100,400])); fl=int(rng.choice([0,1,2,0x40,0x41,0x42,0x80,0xc0,4,5,8,9,0x18,0x19,0x58,0x0a]))
    if kind=="extz2":
        qq,e=F.QE[rng.integers(len(F.QE))]; P=H.make_params(kind,mat,q=qq,e=e,w=w,zdrop=zd,end_bonus=int(rng.choice(F.EB)),flag=fl)
    elif kind=="extd2":
        qq,e,q2,e2=F.DUAL[rng.integers(len(F.DUAL))]; P=H.make_params(kind,mat,q=qq,e=e,q2=q2,e2=e2,w=w,zdrop=zd,end_bonus=int(rng.choice(F.EB)),flag=fl)
    else:
        qq,e,q2,nc=F.SPL[rng.integers(len(F.SPL))]; P=H.make_params(kind,H.simple_mat(5,1,2),q=qq,e=e,q2=q2,noncan=nc,zdrop=zd,flag=int(rng.choice(F.SFLAGS)))
    compare(P,[q],[t],None,int(rng.choice([3,15,24,36])),int(rng.choice([0,2])))
    compare(P,[q],[t],None,-int(rng.choice([16,64,128])),int(rng.choice([0,2])))
    compare(P,[q],[t],None,-(100000+int(rng.choice([16,100,3000]))),int(rng.choice([0,2])))
    if ring_ok(P,[q],[t]): compare(P,[q],[t],None,-200000,int(rng.choice([0,2])))
    n+=1
print("seed",seed,"batches",n,"bad",bad,"secs",round(time.time()-t0),flush=True)
