"""GPU experiment: time every primitive the L-BFGS driver uses at config-2 size."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evcouplings_b200 import synthetic
from evcouplings_b200.engine import CudaEngine
N, L, q = 20000, 200, 21
codes = synthetic.synthetic_msa_codes(N, L, 2)
w = np.random.default_rng(0).uniform(0.1, 1, N).astype(np.float32)
eng = CudaEngine()
p = eng.plm_problem(codes, w, q, -1, 0.01, 39.8)
p.set_x(np.random.default_rng(1).normal(0, 0.05, p.n).astype(np.float32))
def T(name, fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    print("%-28s %8.3f ms" % (name, 1e3 * (time.perf_counter() - t0) / reps))
T("evaluate", lambda: p.evaluate(p.x))
T("evaluate_async", lambda: p.evaluate_async(p.x))
T("dot", lambda: p.dot(p.x, p.g))
T("axpby", lambda: p.axpby(p.d, p.g, -1.0, 0.0))
T("axpby b=1", lambda: p.axpby(p.x, p.d, 1e-9, 1.0))
T("copy", lambda: p.copy(p.xp, p.x))
p.copy(p.gp, p.g); p.axpby(p.x, p.g, -1e-4, 1.0); p.evaluate(p.x)
for s in range(6): p.update_pair(s, p.xp, p.gp)
T("update_pair", lambda: p.update_pair(0, p.xp, p.gp))
T("direction(bound=6)", lambda: p.direction(p.d, 6, 0))
T("norms", lambda: p.norms())
