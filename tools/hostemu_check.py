"""Dev tool: run the host-emulated kernel bodies against the oracle (quick sanity sweep)."""
import ctypes, numpy as np, sys, time
sys.path.insert(0,'/root/repo')
from oracle import poisson_oracle as O
lib = ctypes.CDLL('/root/repo/tests/hostemu/libpsn_hostemu.so')
dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int32)
def P(a): return None if a is None else a.ctypes.data_as(dp)
def PI(a): return None if a is None else a.ctypes.data_as(ip)
def smem(n, theta, sx, sy, rtol=1e-12, maxit=100, nu=2):
    B = sx.shape[0]; u = np.empty((B,n,n)); it = np.zeros(B,np.int32); rr = np.zeros(B); st = np.zeros(B,np.int32)
    rc = lib.hostemu_solve_smem(ctypes.c_int(n), ctypes.c_longlong(B), P(theta), None, P(sx), P(sy), None, None, ctypes.c_int(n), ctypes.c_int(1),
        ctypes.c_double(rtol), ctypes.c_int(maxit), ctypes.c_int(nu), P(u), PI(it), P(rr), PI(st), ctypes.c_int(0), None, None, None, None, None, None)
    assert rc == 0
    return u, it, rr, st
def stream(n, theta, sx, sy, TH, TW, ls, rtol=1e-12, maxit=100, nu=2):
    B = sx.shape[0]; u = np.empty((B,n,n)); it = np.zeros(B,np.int32); rr = np.zeros(B); st = np.zeros(B,np.int32)
    rc = lib.hostemu_solve_stream(ctypes.c_int(n), ctypes.c_longlong(B), P(theta), None, P(sx), P(sy), None, None, ctypes.c_int(n), ctypes.c_int(1),
        ctypes.c_double(rtol), ctypes.c_int(maxit), ctypes.c_int(nu), ctypes.c_int(TH), ctypes.c_int(TW), ctypes.c_int(ls), P(u), PI(it), P(rr), PI(st))
    return u, it, rr, st, rc
def inputs(n, ks, rough, seed=0):
    x = np.linspace(0,1,n); B = len(ks)
    sx = np.stack([np.sin(2*np.pi*k1*x) for k1,k2 in ks]); sy = np.stack([np.sin(2*np.pi*k2*x) for k1,k2 in ks])
    theta = np.random.RandomState(seed).uniform(0.5,2,(B,n,n)) if rough else None
    return sx, sy, theta
if __name__ == "__main__":
    ks = [(0.5,0.7),(2.3,3.7),(11.9,10.3)]
    for n in (5,6,7,13,20,37,40,80):
        for rough in (0,1):
            sx,sy,theta = inputs(n,ks,rough)
            u,it,rr,st = smem(n,theta,sx,sy)
            errs=[]
            for b,(k1,k2) in enumerate(ks):
                f = np.outer(sy[b],sx[b]); th = theta[b] if rough else np.ones((n,n))
                assert np.array_equal(f, O.forcing_term(k1,k2,n))
                errs.append(O.rel_l2(u[b], O.solve_poisson(f,th,n)))
            print(f"smem n={n} rough={rough} iters={it} status={st} relres={rr.max():.1e} err={max(errs):.1e}")
    if len(sys.argv)>1:
      for n,TH,TW,ls in ((96,32,32,1),(100,64,64,1),(160,64,64,1),(160,32,64,2),(131,64,64,1)):
        for rough in (1,):
            sx,sy,theta = inputs(n,ks,rough)
            t=time.time(); u,it,rr,st,rc = stream(n,theta,sx,sy,TH,TW,ls); t=time.time()-t
            errs=[]
            for b,(k1,k2) in enumerate(ks):
                f = np.outer(sy[b],sx[b]); th = theta[b] if rough else np.ones((n,n))
                errs.append(O.rel_l2(u[b], O.solve_poisson(f,th,n)))
            print(f"stream n={n} tile={TH}x{TW} ls={ls} outer={rc} iters={it} status={st} relres={rr.max():.1e} err={max(errs):.1e} ({t:.1f}s)")
