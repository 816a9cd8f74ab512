"""Run every kernel of libevcplm once at small sizes (meant to be run under compute-sanitizer):
    compute-sanitizer --tool memcheck  python scripts/sanitize_all_kernels.py
    compute-sanitizer --tool racecheck python scripts/sanitize_all_kernels.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evcouplings_b200 import synthetic, msa, model_ops, lbfgs
from evcouplings_b200.engine import CudaEngine

eng = CudaEngine()
# Hamming: single-phase (short L) and two-phase (L >= 161) paths, ragged N
for N, L in ((300, 40), (333, 200)):
    codes = synthetic.synthetic_msa_codes(N, L, 1)
    c = eng.hamming_counts(codes, msa.identity_threshold_count(0.8, L))
    assert c.min() >= 1
print("hamming ok")
N, L, q = 300, 24, 21
codes = synthetic.synthetic_msa_codes(N, L, 2)
w = np.random.default_rng(0).uniform(0.1, 1, N).astype(np.float32)
x = np.random.default_rng(1).normal(0, 0.1, L * q + L * (L - 1) // 2 * q * q).astype(np.float32)
ref = None
for fwd, bwd in (("gather", "gather"), ("gather", "tc"), ("tc", "tc"), ("tcfused", "tc")):
    p = eng.plm_problem(codes, w, q, -1, 0.01, 1.0, forward=fwd, backward=bwd, m=3)
    p.set_x(x)
    fx = p.evaluate(p.x)
    g = p.g.cpu().numpy()
    if ref is None:
        ref = (fx, g)
        fi, fij = p.weighted_counts()
        p.fn_scores()
        res = p.fit(np.zeros_like(x), lbfgs.default_params(max_iterations=4))
    else:
        assert abs(fx - ref[0]) < 1e-4 * abs(ref[0]) and np.abs(g - ref[1]).max() < 1e-2
    p.close()
    print("plm", fwd, bwd, "ok", fx)
codes_g = synthetic.to_ignore_gaps_codes(codes)
p = eng.plm_problem(codes_g, w, 20, 20, 0.01, 1.0)
p.set_x(np.zeros(p.n, dtype=np.float32)); p.evaluate(p.x); p.close()
model = dict(L=L, q=q, alphabet=synthetic.ALPHABET, target_seq="A" * L, index_list=np.arange(1, L + 1),
             fi=np.full((L, q), 1.0 / q, dtype=np.float32), h=x[:L * q].reshape(L, q),
             J=x[L * q:].reshape(-1, q, q), fij=np.full((L * (L - 1) // 2, q, q), 1.0 / (q * q), dtype=np.float32))
model_ops.ec_table(model, eng)
model_ops.hamiltonians(model, codes, eng)
print("model ops ok")
