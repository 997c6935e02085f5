"""Short Shor-row relaxation at config 3 (k = 2, 30 x 30, all minors) for the ncu launch list: B nodes, ITERS iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import omc_b200 as omc
from oracle.datagen import config_instance
omc.init(0)
k, A, mask, g = config_instance("C3", 0)
p = omc.Problem(k, A, mask, g, "linear2")
minors, soc = omc.shor_constraint_indexes(p, [1, 2, 3, 4])
p.set_shor(minors, soc)
B = int(os.environ.get("B", "8"))
fr = p.frontier([[]] * B)
ms = fr.relax(omc.default_opts(eps_abs=1e-9, eps_rel=1e-9, max_iter=int(os.environ.get("ITERS", "50"))))
print(f"{B} nodes: {ms:.1f} ms, {fr.stats()}")
