"""GPU diagnostic: tracker state of the batched engine after a few single-step iterations vs oracle/bigblock.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from oracle import bigblock as Bg
from oracle.datagen import config_instance
omc.init(0)
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
k, A, mask, g = config_instance(cfg, 0)
n, m = A.shape
p = omc.Problem(k, A, mask, g, "linear")
for steps in (1, 3):
  for mi in (1, 2, 3):
    o = omc.default_opts(eps_abs=1e-30, eps_rel=1e-30, max_iter=mi, adapt_every=0)
    f = p.frontier([[]], engine="batched"); f.set_tuning(steps_max=steps, steps_start=steps); f.relax(o)
    ro = Bg.solve_relaxation_big(A, mask, g, k, "linear", [], opts=Bg.BigOptions(eps_abs=1e-30, eps_rel=1e-30, max_iter=mi, adaptive_rho=False, steps_max=steps, steps_start=steps))
    st = ro["state"]
    for b, (N, Vn) in enumerate(((n + m, st.V1), (n + k, st.V2), (n, st.V3))):
        V = f.debug_fetch(0, b, N * N).reshape(N, N)
        Z = f.debug_fetch(0, 3 + b, N * 16).reshape(N, 16)
        th = f.debug_fetch(0, 6 + b, 16)
        R = f.debug_fetch(0, 9 + b, N * 16).reshape(N, 16)
        t = st.tr[b]
        pz = t.p
        # compare projectors of the minority part and theta
        Fg = (Z[:, :pz] * np.maximum(th[:pz], 0)) @ Z[:, :pz].T
        Fn = (t.Z * np.maximum(t.th, 0)) @ t.Z.T
        print(f"steps {steps} it {mi} blk {b}: dV {np.abs(V - Vn).max():.2e} dtheta {np.abs(th[:pz] - t.th).max():.2e} dF {np.abs(Fg - Fn).max():.2e} |F| {np.abs(Fn).max():.2e} orthZ {np.abs(Z[:, :pz].T @ Z[:, :pz] - np.eye(pz)).max():.1e}"
              f" theta_gpu {np.round(th[:4], 5)} theta_np {np.round(t.th[:4], 5)} probe-col norm gpu {np.linalg.norm(R[:, pz - 1]):.3e}", flush=True)
    f.close()
p.close()
