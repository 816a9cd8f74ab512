"""GPU experiment: error of each forward/backward implementation against the float64 oracle at full size."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evcouplings_b200 import synthetic
from evcouplings_b200.engine import CudaEngine
from oracle import c_oracle as co

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
L, q = 200, 21
codes = synthetic.synthetic_msa_codes(N, L, 2)
rng = np.random.default_rng(2)
w = rng.uniform(0.05, 1.0, N).astype(np.float32)
n = L * q + L * (L - 1) // 2 * q * q
x = rng.normal(0, 0.05, n).astype(np.float32)
t0 = time.time()
fo, go, _ = co.plm_eval(codes, w.astype(np.float64), x.astype(np.float64), q, 0.0, 0.0, "f64")
print("oracle f64 %.1fs fx=%.6f |g|=%.4e max|g|=%.4e" % (time.time() - t0, fo, np.linalg.norm(go), np.abs(go).max()))
f32, g32, _ = co.plm_eval(codes, w, x, q, 0.0, 0.0, "f32")
print("C fp32 port      : rel L2 %.3e  max abs %.3e  fx rel %.3e" % (
    np.linalg.norm(g32 - go) / np.linalg.norm(go), np.abs(g32 - go).max(), abs(f32 - fo) / abs(fo)))
eng = CudaEngine()
for fwd, bwd in (("gather", "gather"), ("gather", "tc"), ("tc", "tc")):
    p = eng.plm_problem(codes, w, q, -1, 0.0, 0.0, forward=fwd, backward=bwd)
    p.set_x(x)
    fx = p.evaluate(p.x)
    g = p.g.cpu().numpy().astype(np.float64)
    p.close()
    d = g - go
    nh = L * q
    print("fwd=%-6s bwd=%-6s: rel L2 %.3e  max abs %.3e  mean signed err(J) %.3e  fx rel %.3e" % (
        fwd, bwd, np.linalg.norm(d) / np.linalg.norm(go), np.abs(d).max(), d[nh:].mean(), abs(fx - fo) / abs(fo)))
    # error relative to entry magnitude for the large entries
    big = np.abs(go) > 0.1 * np.abs(go).max()
    print("     large entries: median rel err %.3e, signed mean rel err %.3e" % (
        np.median(np.abs(d[big] / go[big])), np.mean(d[big] / go[big])))
