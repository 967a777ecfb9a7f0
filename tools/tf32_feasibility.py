"""CPU emulation of a 3xTF32 (split-operand tensor-core) version of the KG sweep against the reference fixtures.
Round-2 feasibility check: python tools/tf32_feasibility.py kg_small tf32x3   (modes: fp32 | tf32x3 | tf32x1)."""
import numpy as np, sys
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import load_golden, lr_of
case=sys.argv[1]; mode=sys.argv[2]   # mode: fp32 | tf32x3 | tf32x1 | tf32x3trunc
g=load_golden(case)
ns=[int(x) for x in g['ns']]
E=int(g['epochs']); grid=g['grid']
print(case, 'NR', g['out'].shape[0], 'grid', grid.shape, 'epochs', E, 'ns', ns)
f32=np.float32
def rna_tf32(x):
    # round-to-nearest (ties away) to 10 explicit mantissa bits
    b=x.astype(np.float32).view(np.uint32).astype(np.uint64)
    b=(b+0x1000)&0xFFFFE000
    return b.astype(np.uint32).view(np.float32)
def split(x):
    hi=rna_tf32(x); lo=rna_tf32((x-hi).astype(f32)); return hi,lo
def mm(a,b):
    # a [R,K], b [K,P] -> fp32 result
    if mode=='fp32': return (a.astype(np.float64)@b.astype(np.float64)).astype(f32)
    if mode=='tf32x1': return (rna_tf32(a).astype(np.float64)@rna_tf32(b).astype(np.float64)).astype(f32)
    ah,al=split(a); bh,bl=split(b)
    r=ah.astype(np.float64)@bh.astype(np.float64)+ah.astype(np.float64)@bl.astype(np.float64)+al.astype(np.float64)@bh.astype(np.float64)
    return r.astype(f32)
for i,n in enumerate(ns):
    lr=f32(lr_of(g,i))
    P=g['out'][:,:n]; Pxx=g['out_xx'][:,:n]; Ptt=g['out_tt'][:,:n]; PIC=g['out_IC'][:,:n]; PICt=g['out_IC_t'][:,:n]; PBC=g['out_BC'][:,:n]
    NR,NI,NB=P.shape[0],PIC.shape[0],PBC.shape[0]
    facR,facI,facB=f32(2.0/NR),f32(2.0/NI),f32(2.0/NB)
    al=grid[:,0].astype(f32); be=grid[:,1].astype(f32); ga=grid[:,2].astype(f32); ga2=(2.0*grid[:,2]).astype(f32)
    Np=len(grid)
    c=np.full((Np,n),1.0/n,f32)
    # process groups by (alpha,beta)
    keys={}
    for p in range(Np): keys.setdefault((al[p],be[p]),[]).append(p)
    fin=np.zeros(Np,f32); mn=np.full(Np,np.inf,f32)
    for (a,b),idx in keys.items():
        idx=np.array(idx)
        T=((Ptt+a*Pxx).astype(f32)+(b*P).astype(f32)).astype(f32)
        cc=c[idx]
        for j in range(E+1):
            u=mm(P,cc.T); t=mm(T,cc.T)                   # [NR, S]
            d=(t+ga[idx][None,:]*u*u+g['xcos']).astype(f32)
            lossR=np.mean((d-g['f_hat']).astype(np.float64)**2,axis=0)
            first=d; w2=(first*(ga2[idx][None,:]*u)).astype(f32)
            gR=mm(first.T,T)+mm(w2.T,P)                   # [S,n]
            e1=mm(PIC,cc.T)-g['IC_u1']; e2=mm(PICt,cc.T)
            lossI=np.mean(e1.astype(np.float64)**2,axis=0)+np.mean((e2-g['IC_u2']).astype(np.float64)**2,axis=0)
            gI=mm(e1.astype(f32).T,PIC)+mm(e2.astype(f32).T,PICt)
            eb=mm(PBC,cc.T)-g['BC_u']
            lossB=np.mean(eb.astype(np.float64)**2,axis=0)
            gB=mm(eb.astype(f32).T,PBC)
            loss=(lossR+lossI+lossB).astype(f32)
            if j==E: fin[idx]=loss
            else:
                mn[idx]=np.minimum(mn[idx],loss)
                grad=(facR*gR+facI*gI+facB*gB).astype(f32)
                cc=(cc-lr*grad).astype(f32)
        c[idx]=cc
    rf=np.max(np.abs(fin-g[f'n{n}_f'])/np.abs(g[f'n{n}_f']))
    rc=np.max(np.linalg.norm(c-g[f'n{n}_c'],axis=1)/np.linalg.norm(g[f'n{n}_c'],axis=1))
    rm=np.max(np.abs(mn-g[f'n{n}_m'])/np.abs(g[f'n{n}_m']))
    print(f'n={n:2d} mode={mode}: max rel f {rf:.2e}  m {rm:.2e}  c {rc:.2e}   argmax same: {np.argmax(fin)==np.argmax(g[f"n{n}_f"])}')
