"""Scratch (round 2): width of the minority spectral side of the three PSD-block arguments along the exact-projection oracle
trajectory at C4 and at a k=5 analog of C5, cold start.  Decides the panel width of the batched large-block tracker."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import relaxation as R
from oracle.datagen import generate_matrix_completion_data
k, n, m, frac = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
A, mask = generate_matrix_completion_data(k, n, m, int(frac * n * m), 0)
cnt = [0]; log = []
def pp(V):
    lam, Q = np.linalg.eigh(0.5 * (V + V.T))
    b = cnt[0] % 3; it = cnt[0] // 3; cnt[0] += 1
    nrm = np.abs(lam).max()
    log.append((it, b, int((lam > 0).sum()), int((lam < 0).sum()), int((lam > 1e-6 * nrm).sum()), int((lam < -1e-6 * nrm).sum())))
    return (Q * np.maximum(lam, 0)) @ Q.T
R.psd_project = pp
t0 = time.time()
r = R.solve_relaxation(A, mask, 80.0, k, "linear", [], opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=int(sys.argv[5]) if len(sys.argv) > 5 else 6000))
print("iters", r["iters"], "status", r["status"], "obj", r["objective"], "%.0fs" % (time.time() - t0), flush=True)
its = sorted(set([0, 1, 2, 3, 5, 10, 20, 30, 50, 75, 100, 150, 200, 300, 500, 1000, 1500, 2000, 3000, r["iters"] - 1]))
for l in log:
    if l[0] in its:
        print("it %5d blk %d pos %4d neg %4d  pos>1e-6 %4d neg<-1e-6 %4d" % l)
