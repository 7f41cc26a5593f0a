#!/usr/bin/env python
"""numpy emulation of the shared-memory eigensolver (csrc/wct_transform.cu: jacobi_chol_kernel) used to design it on the
CPU: live-channel compaction, pivoted PSD Cholesky, per-sweep refresh, tracked norms, scaled rotations, half-angle
rotation parameters, early stop -- against the first version (Jacobi on the covariance itself).  Prints sweeps,
rotations, residual cosine and the errors of eigenvalues / reconstruction / whitening and colouring matrices vs LAPACK.
Inputs: the pytest spectra and the covariances of the reference-generated golden features (tests/golden/golden_wct.npz,
incl. the rank-deficient cases); pass an .npz of extra covariance matrices as argv[1] if wanted.  No GPU, no oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rr_pairs(rnd, n):
    """circle method, identical to rr_pair() in the kernel source"""
    m = n - 1
    P, Q = [], []
    for k in range(n // 2):
        if k == 0:
            p, q = m, rnd % m
        else:
            p, q = (rnd + k) % m, (rnd - k + m) % m
        if p > q:
            p, q = q, p
        P.append(p)
        Q.append(q)
    return np.array(P), np.array(Q)


def compact(A):
    d = np.diag(A)
    live = [i for i in range(len(d)) if d[i] > 0]
    dead = [i for i in range(len(d)) if d[i] <= 0]
    if len(live) & 1 and dead:
        live.append(dead[0])
    return np.array(live)


def chol_pivot_inplace(S, thr_rel=1e-14):
    k=len(S); S=S.copy(); d=np.diag(S).copy(); done=np.zeros(k,bool)
    thr=thr_rel*d.max(); rank=0
    for j in range(k):
        dm=np.where(done,-np.inf,d); p=int(np.argmax(dm))
        if dm[p]<=thr: break
        l=np.where(done,0.0,S[:,p])/np.sqrt(d[p]); l[p]=np.sqrt(d[p])
        done[p]=True
        nd=~done
        S[np.ix_(nd,nd)]-=np.outer(l[nd],l[nd]); d[nd]-=l[nd]**2
        S[:,p]=l; rank+=1
    S[:,~done]=0.0
    return S,rank
def jacobi_v3(A, tol=1e-10, early=1e-5, maxsw=40, fast=True, track=True, chol=True):
    live=compact(A); S=A[np.ix_(live,live)]; k=len(live)
    if chol:
        G,rank=chol_pivot_inplace(S); lam_is_sq=True; floor2=1e-15*np.trace(S)
    else:
        G=S.copy(); lam_is_sq=False; floor2=(G*G).sum()*1e-30
    s=np.ones(k); si=np.ones(k); stats=[]
    for sweep in range(maxsw):
        G=G*s; s[:]=1; si[:]=1; nrm=(G*G).sum(0)     # refresh
        nrot=0;maxrel=0.0
        for rnd in range(k-1):
            P,Q=rr_pairs(rnd,k)
            x=G[:,P];y=G[:,Q]
            c=(x*y).sum(0)*s[P]*s[Q]
            if track: a=nrm[P];b=nrm[Q]
            else: a=(x*x).sum(0)*s[P]**2;b=(y*y).sum(0)*s[Q]**2
            null=(a<=floor2)|(b<=floor2)
            with np.errstate(all='ignore'):
                skip=(c*c<=tol*tol*a*b)|null
                rel=np.where(null,0,np.abs(c)/np.sqrt(np.abs(a*b)))
                maxrel=max(maxrel,np.nanmax(rel))
                d=b-a;c2=2*c
                h=d*d+c2*c2; r=1/np.sqrt(h)
                cos2=np.abs(d)*r; sin2=np.abs(c2)*r
                cs2=0.5+0.5*cos2; csi=1/np.sqrt(cs2); cs=cs2*csi
                sgn=np.where(d<0,-1.0,1.0)*np.sign(c2)
                sn=0.5*sin2*csi*sgn
                t=sn*csi
            t=np.where(skip,0.0,t); cs=np.where(skip,1.0,cs); sn=np.where(skip,0.0,sn); csi=np.where(skip,1.0,csi)
            if fast:
                tp=t*s[Q]*si[P]; tq=t*s[P]*si[Q]
                G[:,P]=x-tp*y; G[:,Q]=y+tq*x
                s[P]*=cs; s[Q]*=cs; si[P]*=csi; si[Q]*=csi
            else:
                G[:,P]=cs*x-sn*y; G[:,Q]=sn*x+cs*y
            nrm[P]=a-t*c; nrm[Q]=b+t*c
            nrot+=int((~skip).sum())
        stats.append((nrot,maxrel))
        if nrot==0:break
        if early is not None and maxrel<early:break
    G=G*s
    n2=(G*G).sum(0); sig=np.sqrt(n2); lam=n2 if lam_is_sq else sig
    V=np.where(sig>0,G/np.where(sig>0,sig,1),0)
    # residual cosines among non-null columns
    nz=n2>floor2; Vn=V[:,nz]; res=np.abs(Vn.T@Vn-np.eye(nz.sum())).max() if nz.any() else 0
    C=len(A);ev=np.zeros(C);evec=np.zeros((C,C))
    for j in range(k):ev[live[j]]=lam[j];evec[live,live[j]]=V[:,j]
    return ev,evec,stats,res
def check(A,ev,evec):
    def f(ev,evec,pw):
        keep=ev>1e-7*ev.max();e=np.where(keep,ev,1.0)**pw*keep;return (evec*e)@evec.T
    w,v=np.linalg.eigh(A)
    R=(evec*ev)@evec.T
    return (np.abs(np.sort(ev)-np.sort(np.maximum(w,0))).max()/w.max(), np.abs(R-A).max()/np.abs(A).max(),
            np.abs(f(ev,evec,-0.5)-f(w,v,-0.5)).max()/np.abs(f(w,v,-0.5)).max(), np.abs(f(ev,evec,0.5)-f(w,v,0.5)).max()/np.abs(f(w,v,0.5)).max())
if __name__=='__main__':
    mats={}
    if len(sys.argv) > 1:
        Z=np.load(sys.argv[1])
        for k_ in Z.files: mats[k_]=Z[k_]
    for C,rank in [(24,24),(32,20),(64,64),(128,51),(128,128)]:
        g=torch.Generator().manual_seed(C+rank)
        B=torch.randn(C,rank,generator=g,dtype=torch.float64)*torch.logspace(0,-2,rank,dtype=torch.float64)
        A=(B@B.t()).numpy(); mats['t%d.%d'%(C,rank)]=A; mats['t%d.%d+I'%(C,rank)]=A+0.2*np.eye(C)
    W=np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_wct.npz'))
    for nm in ('full_rank','wide','dead_channels','hw_lt_c'):
        for w in ('cF','sF'):
            f=W[nm+'.'+w].astype(np.float64); f=f.reshape(f.shape[-3] if f.ndim>2 else f.shape[0],-1)
            fc=f-f.mean(1,keepdims=True); mats[nm+'.'+w]=fc@fc.T/(fc.shape[1]-1)
    for key,A in mats.items():
        for name,kw in [('old',dict(chol=False,fast=False,track=False,early=None)),('v3',dict()),('v3 e3e-6',dict(early=3e-6))]:
            ev,evec,st,res=jacobi_v3(A,**kw)
            print('%-16s %-9s sw=%2d rots=%6d res=%.1e  ev=%.1e recon=%.1e W=%.1e Col=%.1e'%(key,name,len(st),sum(s[0] for s in st),res,*check(A,ev,evec)), '%.0e'%st[-1][1])
