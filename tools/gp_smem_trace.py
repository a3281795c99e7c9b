"""Phase stamps of the shared-memory GP kernel (variant built with -DRVLP_GPS_TRACE): one sample on CTA 0.
RVLP_LIB=build_variants/lib_gps_trace.so python tools/gp_smem_trace.py [N] [S]"""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
os.environ["RVLP_GP_KERNEL"] = "smem"
from ravest_b200 import fit, workloads
N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
S = int(sys.argv[2]) if len(sys.argv) > 2 else 148
spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N, seed=505)
theta[:, -1] = np.abs(theta[:, -1])
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device="cuda")
NT = (N + 7) // 8
buf = torch.zeros(S + NT * 4 * 6, dtype=torch.float64, device="cuda")
for _ in range(3): post.ctx.logprob(th, out=buf[:S])
torch.cuda.synchronize()
st = buf[S:].cpu().numpy().reshape(NT, 4, 6)
print("workers: start | last, ahead, wait L_jj, solve, wait column   diag (w4): start | last, factor, ahead, -, wait column")
for j in range(NT):
    row = []
    for w in range(4):
        a = st[j, w]
        row.append("w%d %6d |%5d %5d %5d %5d %5d" % (w, a[0], a[1] - a[0], a[2] - a[1], a[3] - a[2], a[4] - a[3], a[5] - a[4]))
    print("j=%2d " % j + "  ".join(row))
