"""Round-2 idea for a cheap infeasibility test: a feasible node has p* <= c0 = 1/2 ||P_Omega(A)||^2
(X = 0, Theta = 0 with any feasible (Y, U)), so a certified lower bound above c0 proves infeasibility without the
d mu pass.  This script instruments a copy of the oracle's ADMM loop and prints, for the infeasible linear3 chain of
tests (certificate at iteration 5450), the dual objective and the certified bound obj_d - ||r_d||_inf ||w||_1 at
every 250th iteration, next to c0."""
import inspect
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import oracle.relaxation as R                      # noqa: E402
from oracle.datagen import generate_matrix_completion_data   # noqa: E402
from conftest import feasible_chain                # noqa: E402

code = inspect.getsource(R.solve_relaxation)
hook = '''            res_p, res_d = rp, rd
            if it % 250 == 0:
                dual_ = (-0.5 * float(np.sum(Mk * st.X * st.X)) + c.c0
                         + np.trace(st.m2[n:, n:]) + c.a * np.trace(st.m3) + c.ktr * st.m4 + float(c.beta @ st.mg)
                         - float(np.sum(np.where(st.m5 < 0, st.m5 * c.lo, st.m5 * c.hi)))
                         - float(np.sum(np.where(st.mv < 0, st.mv * c.lb, st.mv * c.ub))))
                trTb = c.c0 / c.cT
                w1 = n * c.ktr + np.sqrt(n * m * c.ktr * trTb) + m * trTb + n * c.k * c.sa
                print("it %5d  dual %.4e  bound %.4e  c0 %.4e  rd %.2e" % (it, dual_, dual_ - rd * w1, c.c0, rd), flush=True)
'''
assert "            res_p, res_d = rp, rd\n" in code
code = code.replace("            res_p, res_d = rp, rd\n", hook, 1)
ns = dict(R.__dict__)
exec(code, ns)

n, m, k, ct, L = 6, 9, 2, "linear3", 10
rng = np.random.default_rng(100 * n + 10 * k + L)
A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.6 * n * m)), 5)
cuts = feasible_chain(ct, n, k, L, rng)
r = ns["solve_relaxation"](A, mask, 20.0, k, ct, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=6000))
print("status", r["status"], "iters", r["iters"])
