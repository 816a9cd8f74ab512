"""GPU experiment: wall time per L-BFGS iteration of a full run_plmc at BASELINE config 2 (N=50k, L=200)."""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evcouplings_b200 import synthetic, tools
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 200
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
codes = synthetic.synthetic_msa_codes(N, L, 2)
d = tempfile.mkdtemp()
a2m = os.path.join(d, "a.a2m")
t0 = time.time(); synthetic.write_a2m(a2m, codes); print("write a2m %.1fs" % (time.time() - t0))
t0 = time.time()
res, run = tools.run_plmc(a2m, os.path.join(d, "o_ECs.txt"), os.path.join(d, "o.model"), focus_seq="seq0", theta=0.8,
                          iterations=iters, lambda_h=0.01, lambda_J=0.01 * 20 * (L - 1), return_run=True)
print("total %.2fs" % (time.time() - t0), run.timings)
print("status", res.optimization_status, "iterations", run.lbfgs.iterations, "evaluations", run.lbfgs.evaluations,
      "ms/iteration %.2f" % (1e3 * run.timings["optimisation_s"] / max(1, run.lbfgs.iterations)),
      "ms/evaluation %.2f" % (1e3 * run.timings["optimisation_s"] / max(1, run.lbfgs.evaluations)))
print(res.iteration_table.head(4)); print(res.iteration_table.tail(3))
