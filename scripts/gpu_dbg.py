import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
from omc_b200 import Problem, Cut, default_opts
from oracle.datagen import config_instance
omc_b200.init(0)
k, A, mask, g = config_instance("C1", 0)
p = Problem(k, A, mask, g, "linear")
rng = np.random.default_rng(5)
x = rng.standard_normal(10); x /= np.linalg.norm(x); Uh = 0.3 * rng.standard_normal((10, 1))
cid = p.add_cut(x, Uh)
opts = default_opts(eps_abs=1e-30, eps_rel=1e-30, max_iter=3, adapt_every=0, check_every=1)
f = p.frontier([[Cut(cid, x, Uh, ["right"])]]); f.relax(opts); print(f.fetch(False)[0]); print(f.profile()[0, 8:15])
