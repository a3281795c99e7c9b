import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import ravest_b200
from ravest_b200 import fit, workloads
from oracle import oracle_c
for npl,N,S in ((1,120,10000),(2,57,10000),(1,170,4000),(1,30,20000),(1,200,2000)):
    spec, theta = workloads.make_c5(n_samples=S, n_planets=npl, n_epochs=N)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device='cuda'); out = torch.empty(S, dtype=torch.float64, device='cuda')
    for _ in range(2): post.ctx.logprob(th, out=out)
    torch.cuda.synchronize(); a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): post.ctx.logprob(th, out=out)
    b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b)/5
    ref = oracle_c.OracleProblem(spec).logprob(theta[:200]); got = out[:200].cpu().numpy()
    fin = np.isfinite(ref)
    err = np.max(np.abs(got[fin]-ref[fin])/(1e-7+1e-11*np.abs(ref[fin])))
    print(f"npl={npl} N={N} S={S}: {ms:.3f} ms  {S/ms*1e3:.3e} logprob/s  flops~{S*(N**3/3+N*N*50)/ms/1e9:.2f} TFLOP/s  err/tol={err:.3g} infmatch={np.array_equal(np.isneginf(got),np.isneginf(ref))}")
