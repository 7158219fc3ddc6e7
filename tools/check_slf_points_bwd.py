"""Derivation check of nrc_slf_points_bwd WITHOUT a GPU: a float64 NumPy transcription of the kernel's backward formulas
(csrc/slf.cu: softmax VJP, env-alpha path, fold sign, d t / d s of the power-ladder inverse, contraction VJP) against autograd
through the oracle's predict_points.  Usage (repo root): python tools/check_slf_points_bwd.py   -> relative L2 errors ~1e-7."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import surface_light_field as oslf
from tests.util import f32, gen
def sigmoid(x): return 1/(1+np.exp(-x))
def kernel_bwd(raw,o,v,n,dn,df,near,far,warp,ups,c=2.0):
    # numpy mirror of slf_points_bwd_kernel (float64)
    P,W=raw.shape; g=np.zeros_like(raw)
    gp,gw,gsd,gd,ge=ups
    p_,pre=warp if warp else (1.0,1.0)
    def plf(x):
        x=x*pre; xs=abs(x)/abs(p_-1); return np.sign(x)*(abs(p_-1)/p_*((xs+1)**p_-1))
    def pli(y):
        ratio=p_/abs(p_-1); return np.sign(y)*(abs(p_-1)*((ratio*abs(y)+1)**(1/p_)-1))/pre
    s_near,s_far=(plf(dn),plf(df)) if warp else (dn,df)
    for r in range(P):
        row=raw[r]; ea=sigmoid(row[W-1]+2.0)
        rw=row[4:8*n:8]; sm=np.exp(rw-rw.max()); sm/=sm.sum()
        start=np.linspace(1e-8,1-1e-8,n)
        S=[];
        for i in range(n):
            o0,o1=row[8*i],row[8*i+1]; sig=sigmoid(o1-2.0); off=o0*1.0/n*sig
            sp=off+start[i]; fl=np.floor(sp); frac=sp-fl; even=(int(fl)%2)==0
            s=frac if even else 1-frac; sgn=1.0 if even else -1.0
            u=s*s_far+(1-s)*s_near; t=pli(u) if warp else u
            mask=float(t>dn and t<df and t>near and t<far)
            dtds=0.0
            if t>dn and t<df:
                dtdu=1.0
                if warp:
                    ratio=p_/abs(p_-1); dtdu=(ratio*abs(u)+1)**(1/p_-1)/pre
                dtds=dtdu*(s_far-s_near)
            t=min(max(t,dn),df)
            S.append((s,t,mask,sgn,dtds,sig,o0))
        dot=0; g_ea=ge[r,3]
        for i in range(n):
            s,t,mask,sgn,dtds,sig,o0=S[i]
            g_sm=gw[r,i]*mask*ea+gsd[r,0]*s; dot+=sm[i]*g_sm; g_ea+=gw[r,i]*sm[i]*mask
        for i in range(n):
            s,t,mask,sgn,dtds,sig,o0=S[i]
            g_sm=gw[r,i]*mask*ea+gsd[r,0]*s
            g[r,8*i+4]=sm[i]*(g_sm-dot)
            x=o[r]+t*v[r]; y=x/c; m=(y*y).sum()
            gg=gp[r,i]
            if m<=1: a=gg
            else:
                rr=np.sqrt(m); sc=(2*rr-1)/m; dsdm=1/(rr*m)-(2*rr-1)/m**2; gy=(gg*y).sum(); a=sc*gg+2*dsdm*gy*y
            a=a/c
            g_t=gd[r,i]+(a*v[r]).sum()
            g_s=gsd[r,0]*sm[i]+g_t*dtds
            g_off=g_s*sgn; k=1.0/n
            g[r,8*i]=g_off*k*sig; g[r,8*i+1]=g_off*o0*k*sig*(1-sig)
        for a_ in range(3):
            x=row[W-4+a_]-2.0; g[r,W-4+a_]=ge[r,a_]*sigmoid(x)
        g[r,W-1]=g_ea*ea*(1-ea)
    return g
for n,warp in ((8,(-1.5,2.0)),(4,None),(1,(-1.5,2.0))):
    g=gen(5100+n); P,W=333,8*n+4
    raw=f32(g.normal(size=(P,W))*1.5); o=f32(g.normal(size=(P,3))*1.2); v=f32(g.normal(size=(P,3))); v=v/v.norm(dim=-1,keepdim=True)
    ups=[f32(g.normal(size=s)) for s in ((P,n,3),(P,n),(P,1),(P,n),(P,4))]
    rw=raw.clone().requires_grad_(True)
    pp=oslf.predict_points(rw,o,v,n,5e-2,2.0,0.1,1.7,raydist=warp)
    loss=sum((a*b).sum() for a,b in zip((pp["points"],pp["ref_weights"],pp["s_dist"],pp["distances"],pp["env_rgba"]),ups))
    loss.backward()
    gk=kernel_bwd(raw.double().numpy(),o.double().numpy(),v.double().numpy(),n,5e-2,2.0,0.1,1.7,warp,[u.double().numpy() for u in ups])
    ref=rw.grad.double().numpy()
    print(n,warp,np.linalg.norm(gk-ref)/np.linalg.norm(ref), np.abs(gk-ref).max(), 'mask frac', float(pp['ref_mask'].mean()))
