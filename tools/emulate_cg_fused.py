"""Numpy emulation of csrc/cg_fused.cu's data flow (windows per block, staged column ranges with clamping,
three phases separated by grid barriers) against a plain PCG on the same BC-modified matrix.
CPU only; checks the indexing and phase order of the kernel, not the CUDA code itself.
    python tools/emulate_cg_fused.py      # same iteration count and 1e-16 agreement for 1, 2, 3, 148 "SMs"
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'nmpde-wave-equation_b200'))
import numpy as np
from oracle import oracle as O
from wavegpu.problems import problem
KW, KS = 1024, 32
def build_sell(rowptr, col, val, n):
    nslots = (n + KW - 1)//KW*KW
    rowlen = np.diff(rowptr)
    row_of = -np.ones(nslots, dtype=np.int64)
    for w in range(nslots//KW):
        rows = np.arange(w*KW, min((w+1)*KW, n))
        order = rows[np.argsort(-rowlen[rows], kind='stable')]
        row_of[w*KW:w*KW+len(order)] = order
    nsl = nslots//KS
    slice_len = np.zeros(nsl, dtype=np.int64)
    for s in range(nsl):
        r = row_of[s*KS:(s+1)*KS]; r = r[r>=0]
        slice_len[s] = rowlen[r].max() if len(r) else 0
    slice_ptr = np.concatenate([[0], np.cumsum(slice_len*KS)])
    pc = np.zeros(slice_ptr[-1], dtype=np.int64)   # padding col = own_off = 0
    pv = np.zeros(slice_ptr[-1])
    for s in range(nsl):
        for lane in range(KS):
            r = row_of[s*KS+lane]
            if r < 0: continue
            for k in range(rowlen[r]):
                pc[slice_ptr[s]+KS*k+lane] = col[rowptr[r]+k]; pv[slice_ptr[s]+KS*k+lane] = val[rowptr[r]+k]
    return row_of, slice_ptr, pc, pv, rowlen
def fused_cg(row_of, slice_ptr, pc, pv, rowlen, rowptr, dinv, x, b, sms, reduce=1e-6, tol=1e-12, maxit=10000):
    n = len(x); nwin = len(row_of)//KW
    wpb = (nwin + sms - 1)//sms; grid = (nwin + wpb - 1)//wpb
    # plan: per-window col range over real entries, per block union
    c0 = np.zeros(grid, dtype=np.int64); cn = np.zeros(grid, dtype=np.int64)
    for bk in range(grid):
        lo, hi = 1<<60, -1
        for w in range(bk*wpb, min(nwin,(bk+1)*wpb)):
            for t in range(KW):
                r = row_of[w*KW+t]
                if r < 0: continue
                s = (w*KW+t)//KS; lane = t % KS
                for k in range(rowlen[r]):
                    c = pc[slice_ptr[s]+KS*k+lane]; lo=min(lo,c); hi=max(hi,c)
        c0[bk]=lo; cn[bk]=hi-lo+1
    # start state (k_spmv residual epilogue + k_cg_start)
    A = lambda v: np.array([sum(pv_csr[rowptr[i]+k]*v[col_csr[rowptr[i]+k]] for k in range(rowlen[i])) for i in range(n)])
    g = Amat @ x - b; h = dinv*g; d = -h
    gg = g@g; gh_old = gh_new = g@h; res0 = np.sqrt(gg); red = res0*reduce; it = 0
    status = 1 if (res0 <= red or res0 <= tol) else 0
    if status: return x, it
    sx = {}; sg = {}; hk = {}
    for bk in range(grid):
        for k in range(wpb):
            w = bk*wpb+k
            if w >= nwin: continue
            for t in range(KW):
                r = row_of[w*KW+t]
                if r >= 0: sx[(bk,k,t)] = x[r]; sg[(bk,k,t)] = g[r]
    while True:
        # stage + phase 1 (all blocks), then barrier
        part = np.zeros(grid); sds = []
        for bk in range(grid):
            sd = d[c0[bk]:c0[bk]+cn[bk]].copy(); sds.append(sd)
            for k in range(wpb):
                w = bk*wpb+k
                if w >= nwin: continue
                for t in range(KW):
                    s_ = (w*KW+t)//KS; lane = t%KS; r = row_of[w*KW+t]
                    ln = (slice_ptr[s_+1]-slice_ptr[s_])//KS
                    acc = 0.0
                    for j in range(ln):
                        q = slice_ptr[s_]+KS*j+lane
                        idx = min((pc[q]-c0[bk]) % (1<<32), cn[bk]-1)
                        acc = acc + pv[q]*sd[idx]
                    hk[(bk,k,t)] = acc
                    if r >= 0: part[bk] += acc*sd[r-c0[bk]]
        dAd = part.sum(); alpha = gh_old/dAd
        p2 = np.zeros((grid,2))
        for key in sx:
            bk,k,t = key; r = row_of[(bk*wpb+k)*KW+t]
            gi = sg[key] + alpha*hk[key]; sg[key]=gi; p2[bk,0]+=gi*gi
            hi_ = dinv[r]*gi; hk[key]=hi_; p2[bk,1]+=gi*hi_
        gg, gh_new = p2[:,0].sum(), p2[:,1].sum()
        res = np.sqrt(abs(gg)); it += 1
        if res <= red or res <= tol: status = 1
        elif it >= maxit or np.isnan(res): status = 2
        beta = gh_new/gh_old
        newd = d.copy()
        for key in sx:
            bk,k,t = key; r = row_of[(bk*wpb+k)*KW+t]
            dold = sds[bk][r-c0[bk]]
            sx[key] += alpha*dold
            if status == 0: newd[r] = beta*dold - hk[key]
        d = newd; gh_old = gh_new
        if status: break
    xo = x.copy()
    for key,v in sx.items():
        bk,k,t = key; xo[row_of[(bk*wpb+k)*KW+t]] = v
    return xo, it
p = problem("standing-mode-wsol", Nel="70, 40", R=1, Dt="0.05")
o = O.Oracle.from_params(p); o.newmark_init(0.05,0.25,0.5)
rowptr, col = o.csr(); val = o.values(O.Oracle.M) + 0.25*0.05**2*o.values(O.Oracle.K)
n = o.n
import scipy.sparse as sp
Amat = sp.csr_matrix((val, col, rowptr), shape=(n,n))
# Dirichlet rows -> d0 e_i
bd = o.boundary_dofs(); d0 = abs(Amat[0,0])
Al = Amat.tolil()
for i in bd: Al.rows[i] = [i]; Al.data[i] = [d0]
# keep the pattern: rebuild values on the original pattern
val2 = val.copy()
for i in bd:
    for e in range(rowptr[i], rowptr[i+1]): val2[e] = d0 if col[e]==i else 0.0
Amat = sp.csr_matrix((val2, col, rowptr), shape=(n,n))
pv_csr, col_csr = val2, col
rng = np.random.default_rng(3); b = rng.standard_normal(n); b[bd] = 0.0
x0 = np.zeros(n)
dinv = 1.0/Amat.diagonal()
row_of, slice_ptr, pc, pv, rowlen = build_sell(rowptr, col, val2, n)
print('n',n,'nwin',len(row_of)//KW)
for sms in (1,2,3,148):
    x, it = fused_cg(row_of, slice_ptr, pc, pv, rowlen, rowptr, dinv, x0.copy(), b, sms)
    # reference: plain PCG (deal.II form)
    xr = x0.copy(); g = Amat@xr-b; h=dinv*g; d=-h; gh=g@h; res0=np.sqrt(g@g); itr=0
    while True:
        Ad = Amat@d; alpha=gh/(d@Ad); xr+=alpha*d; g+=alpha*Ad; res=np.sqrt(g@g); itr+=1
        if res<=res0*1e-6 or res<=1e-12: break
        h=dinv*g; ghn=g@h; beta=ghn/gh; d=beta*d-h; gh=ghn
    print('sms',sms,'its',it,itr,'relerr',np.abs(x-xr).max()/np.abs(xr).max(), 'residual', np.linalg.norm(Amat@x-b)/np.linalg.norm(b))
