"""
TEST-ONLY engine: plugs the CPU oracle behind the host logic of evcouplings_b200.tools.run_plmc so
that the file writers, the log, PlmcResult and the L-BFGS control flow can be exercised without a
GPU (and so that the reference's own couplings protocol can be run end-to-end over our boundary in
this container).  Lives under tests/ on purpose: the product has no CPU path.
"""
import numpy as np

from evcouplings_b200 import lbfgs
from oracle import c_oracle as co
from oracle import plm_oracle as po


class OracleProblem(object):
    rank = 0

    def __init__(self, codes, weights, q, gap_code, lambda_h, lambda_J, m=6, precision="f64"):
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.w = np.asarray(weights, dtype=np.float64)
        self.L, self.q, self.gap_code = codes.shape[1], q, gap_code
        self.lambda_h, self.lambda_J = lambda_h, lambda_J
        self.n = self.L * q + self.L * (self.L - 1) // 2 * q * q
        self.m = m
        self.precision = precision
        z = lambda: np.zeros(self.n)
        self.x, self.g, self.xp, self.gp, self.d = z(), z(), z(), z(), z()
        self.S = np.zeros((m, self.n))
        self.Y = np.zeros((m, self.n))
        self.ys = np.zeros(m)
        self.yy = 0.0
        self.last_negloglk = float("nan")
        self.evaluations = 0

    def evaluate(self, x):
        fx, g, nll = co.plm_eval(self.codes, self.w, x, self.q, self.lambda_h, self.lambda_J,
                                 precision=self.precision)
        self.g[:] = g
        self.last_negloglk = nll
        self.evaluations += 1
        return fx

    def dot(self, a, b):
        return float(np.dot(a, b))

    def copy(self, dst, src):
        dst[:] = src

    def axpby(self, y, x, a, b):
        y[:] = a * x if b == 0.0 else a * x + b * y

    def update_pair(self, slot, xp, gp):
        self.S[slot] = self.x - xp
        self.Y[slot] = self.g - gp
        self.ys[slot] = float(np.dot(self.Y[slot], self.S[slot]))
        self.yy = float(np.dot(self.Y[slot], self.Y[slot]))

    def direction(self, d, bound, end):
        m = self.m
        d[:] = -self.g
        alpha = np.zeros(m)
        j = end
        for _ in range(bound):
            j = (j + m - 1) % m
            alpha[j] = np.dot(self.S[j], d) / self.ys[j]
            d -= alpha[j] * self.Y[j]
        last = (end + m - 1) % m
        d *= self.ys[last] / self.yy
        for _ in range(bound):
            beta = np.dot(self.Y[j], d) / self.ys[j]
            d += (alpha[j] - beta) * self.S[j]
            j = (j + 1) % m

    def weighted_counts(self):
        X = po.one_hot(self.codes, self.q)
        Xw = X * self.w[:, None, None]
        fi = Xw.sum(axis=0)
        F = np.einsum("nia,njb->ijab", Xw, X, optimize=True)
        iu, ju = np.triu_indices(self.L, 1)
        return fi, F[iu, ju]

    def fn_scores(self):
        Jt = self.x[self.L * self.q:].reshape(-1, self.q, self.q)
        return np.sqrt((Jt ** 2).sum(axis=(1, 2)))

    def get_x(self):
        return self.x.astype(np.float32)

    def norms(self):
        nh = self.L * self.q
        return float(np.linalg.norm(self.x[:nh])), float(np.linalg.norm(self.x[nh:]))

    def fit(self, x0, params, progress=None):
        self.x[:] = x0
        return lbfgs.minimize(self, params, progress)

    def close(self):
        pass


class OracleEngine(object):
    rank = 0
    world = 1

    def __init__(self, precision="f64"):
        self.precision = precision

    def hamming_counts(self, codes, min_identical):
        return co.hamming_counts(codes, min_identical)

    def plm_problem(self, codes, weights, q, gap_code, lambda_h, lambda_J, m=6):
        return OracleProblem(codes, weights, q, gap_code, lambda_h, lambda_J, m, self.precision)


class ShardedOracleProblem(OracleProblem):
    """Data-parallel twin of the CUDA problem for the gloo tests of the multi-rank plumbing: this rank's
    contiguous block of sequences, ONE all-reduce of [g, -loglk] per evaluation, regulariser after."""

    def __init__(self, engine, *a, **k):
        super().__init__(*a, **k)
        from evcouplings_b200.dist import shard_bounds
        self.engine = engine
        lo, hi = shard_bounds(self.codes.shape[0], engine.world, engine.rank)
        self.c_loc, self.w_loc = self.codes[lo:hi], self.w[lo:hi]

    def evaluate(self, x):
        import torch
        nll, g, _ = co.plm_eval(self.c_loc, self.w_loc, x, self.q, 0.0, 0.0, precision="f64")
        packed = torch.from_numpy(np.concatenate([g, [nll]]))
        self.engine.coll.all_reduce_sum(packed)
        packed = packed.numpy()
        nh = self.L * self.q
        lam = np.concatenate([np.full(nh, self.lambda_h), np.full(self.n - nh, self.lambda_J)])
        self.g[:] = packed[:-1] + 2 * lam * x
        self.last_negloglk = float(packed[-1])
        self.evaluations += 1
        return self.last_negloglk + float((lam * x * x).sum())

    def weighted_counts(self):
        import torch
        X = po.one_hot(self.c_loc, self.q)
        Xw = X * self.w_loc[:, None, None]
        fi = Xw.sum(axis=0)
        F = np.einsum("nia,njb->ijab", Xw, X, optimize=True)
        iu, ju = np.triu_indices(self.L, 1)
        tf, tF = torch.from_numpy(np.ascontiguousarray(fi)), torch.from_numpy(np.ascontiguousarray(F[iu, ju]))
        self.engine.coll.all_reduce_sum(tf)
        self.engine.coll.all_reduce_sum(tF)
        return tf.numpy(), tF.numpy()


class ShardedOracleEngine(object):
    """OracleEngine as one rank of an initialised torch.distributed (gloo) group."""

    def __init__(self, precision="f64"):
        from evcouplings_b200.dist import Collective
        self.coll = Collective()
        self.rank, self.world = self.coll.rank, self.coll.world
        self.precision = precision

    def hamming_counts(self, codes, min_identical):
        return co.hamming_counts(codes, min_identical)

    def plm_problem(self, codes, weights, q, gap_code, lambda_h, lambda_J, m=6):
        return ShardedOracleProblem(self, codes, weights, q, gap_code, lambda_h, lambda_J, m, self.precision)

    def agree_any(self, flag):
        import torch
        t = torch.tensor([1 if flag else 0], dtype=torch.int32)
        self.coll.all_reduce_max(t)
        return bool(int(t.item()))
