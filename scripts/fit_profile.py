"""GPU experiment: host-side time per L-BFGS primitive inside a real fit (config 2)."""
import os, sys, time, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evcouplings_b200 import synthetic, lbfgs
from evcouplings_b200.engine import CudaEngine, CudaPlmProblem
N, L, q = 50000, 200, 21
codes = synthetic.synthetic_msa_codes(N, L, 2)
w = np.random.default_rng(0).uniform(0.1, 1, N).astype(np.float32)
eng = CudaEngine()
p = eng.plm_problem(codes, w, q, -1, 0.01, 39.8)
acc = collections.defaultdict(lambda: [0, 0.0])
def wrap(name):
    fn = getattr(p, name)
    def inner(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        acc[name][0] += 1; acc[name][1] += dt
        return r
    setattr(p, name, inner)
for nme in ("evaluate", "dot", "copy", "axpby", "update_pair", "direction", "norms"):
    wrap(nme)
x0 = np.zeros(p.n, dtype=np.float32)
t0 = time.perf_counter()
res = p.fit(x0, lbfgs.default_params(max_iterations=30), lambda k, fx, xn, gn, st, nls: p.norms() and False)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(res, "total %.1f ms" % (1e3 * tot))
s = 0
for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-12s calls %4d  total %8.2f ms  per call %7.3f ms" % (k, c, 1e3 * t, 1e3 * t / c)); s += t
print("accounted %.1f ms; python/other %.1f ms" % (1e3 * s, 1e3 * (tot - s)))
