"""TEST INFRASTRUCTURE ONLY — numpy twin of the GPU PDHG for the fir_qp_cvx SOCP (fir_qp_cvx.m:145-166) in the
eliminated form the product solves: minimise ||x|| + obj*max_i ||(x_i, x_{n+i})|| s.t. one disk per grid point.
Blocks: disk pairs (dual prox = v - sigma*P_disk(v/sigma)), group block (projection onto the l1,2 ball of radius
obj), primal norm term (block soft-threshold).  Parity: UNPINNED by the reference (no CVX); pinned against SciPy
trust-constr on small designs (tests/golden/fir_qp_known.json)."""
import time

import numpy as np

from .fir_problems import build_fir_qp, objective_fir_qp, violation_fir_qp  # noqa: F401

def solve_qp_twin(p, max_iter=100000, check=64, eps_pr=8e-7, eps_gap=5e-5, eps_dr=1e-4, verbose=False):
    n=p['n']; w=p['w']; m=w.size; kk=np.arange(n)
    Ew=np.exp(-1j*np.outer(w,kk)); Ar=np.hstack([Ew.real,-Ew.imag]); Ai=np.hstack([Ew.imag,Ew.real])
    N=2*n
    K=np.zeros((2*m+2*n,N)); K[0:2*m:2]=Ar; K[1:2*m:2]=Ai
    # group block rows: pair k -> rows 2m+2k (x_k), 2m+2k+1 (x_{n+k})
    for k in range(n):
        K[2*m+2*k,k]=1; K[2*m+2*k+1,n+k]=1
    cs=np.sqrt((K**2).sum(0)); s=cs.mean(); assert np.allclose(cs,s)
    K=K/s; lam=1.0/s       # objective ||x|| = ||z||/s ;  obj*max||pair(x)|| = (obj/s) max||pair(z)||
    wgt=p['obj']/s
    cr=np.empty(2*m); cr[0::2]=p['center'].real; cr[1::2]=p['center'].imag
    R=p['radius']
    v=np.random.default_rng(0).normal(size=N)
    for _ in range(60):
        v=K.T@(K@v); nk=np.linalg.norm(v); v/=nk
    eta=0.9/np.sqrt(nk); omega=1.0
    z=np.zeros(N); y=np.zeros(2*m+2*n); zs=np.zeros(N); ys=np.zeros_like(y); cnt=0; z0=z.copy(); y0=y.copy()
    last=np.inf; prev=np.inf; since=0
    def proxz(v,t):
        nv=np.linalg.norm(v); return v*max(0.0,1-t*lam/nv) if nv>0 else v
    def proj_disks(wv):   # wv (2m,) -> projected onto disks
        d=wv-cr; dr_=np.hypot(d[0::2],d[1::2]); sc=np.where(dr_>R, R/np.maximum(dr_,1e-300), 1.0)
        out=cr.copy(); out[0::2]+=d[0::2]*sc; out[1::2]+=d[1::2]*sc; return out
    def proj_l12(u,rad):  # u (2n,) pairs ; ball sum ||u_k|| <= rad
        r=np.hypot(u[0::2],u[1::2])
        if r.sum()<=rad: return u
        srt=-np.sort(-r); css=np.cumsum(srt)-rad; kidx=np.arange(1,n+1); cond=srt-css/kidx>0; rr=np.nonzero(cond)[0][-1]; th=css[rr]/(rr+1)
        sc=np.maximum(r-th,0)/np.maximum(r,1e-300); out=u.copy(); out[0::2]*=sc; out[1::2]*=sc; return out
    def metrics(zz,yy):
        Kz=K@zz; d=Kz[:2*m]-cr; pr=max(0.0,(np.hypot(d[0::2],d[1::2])-R).max())
        g=K.T@yy
        dr=np.abs(zz-proxz(zz-g,1.0)).max()
        pk=np.hypot(Kz[2*m::2],Kz[2*m+1::2]).max()
        pobj=lam*np.linalg.norm(zz)+wgt*pk*s   # careful: group rows = z/s pairs -> pk in x units*? K rows identity/s: Kz = z/s = x ; max||pair(x)||*obj
        pobj=lam*np.linalg.norm(zz)+p['obj']*pk
        yd=yy[:2*m]; hs=(cr*yd).sum()+(R*np.hypot(yd[0::2],yd[1::2])).sum()
        dobj=-hs+g@zz+lam*np.linalg.norm(zz)
        return pr,dr,pobj,dobj
    t0=time.time()
    for it in range(1,max_iter+1):
        tau=eta/omega; sig=eta*omega
        zn=proxz(z-tau*(K.T@y),tau); zb=2*zn-z
        vv=y+sig*(K@zb)
        yn=np.empty_like(y)
        yn[:2*m]=vv[:2*m]-sig*proj_disks(vv[:2*m]/sig)
        # group block: f = obj*max||pair||  conj = indicator of l1,2 ball radius obj  -> y+ = Proj_ball(v)
        yn[2*m:]=proj_l12(vv[2*m:],p['obj'])
        # exact zeros for inside-disk
        z,y=zn,yn; zs+=z; ys+=y; cnt+=1; since+=1
        if it%check: continue
        za,ya=zs/cnt,ys/cnt
        ma=metrics(za,ya); mc=metrics(z,y)
        ea=max(ma[0],ma[1],abs(ma[2]-ma[3])); ec=max(mc[0],mc[1],abs(mc[2]-mc[3]))
        if ea<ec: cz,cy,ce,cm=za,ya,ea,ma
        else: cz,cy,ce,cm=z,y,ec,mc
        if verbose and it%(check*32)==0: print(it,'pr %.1e dr %.1e obj %.7f d %.7f om %.2e'%(cm[0],cm[1],cm[2],cm[3],omega),'%.1fs'%(time.time()-t0))
        if cm[0]<=eps_pr and cm[1]<=eps_dr and abs(cm[2]-cm[3])<=eps_gap*abs(cm[2]):
            return cz/s, cm, it
        if ce<=0.2*last or (ce<=0.8*last and ce>prev) or since>=0.36*it:
            dz=np.linalg.norm(cz-z0); dy=np.linalg.norm(cy-y0)
            if dz>1e-12 and dy>1e-12: omega=np.exp(0.5*np.log(dy/dz)+0.5*np.log(omega))
            z,y=cz.copy(),cy.copy(); z0=z.copy(); y0=y.copy(); zs[:]=0; ys[:]=0; cnt=0; last=ce; since=0
        prev=ce
    return z/s, mc, max_iter

