"""
CUDA engine: device memory / streams / collectives (torch) around the C ABI of
libevcplm.so.  This is the only execution path of the package -- there is no
CPU implementation behind it; construction raises EngineUnavailableError when
the library or a CUDA device is missing.

Multi-GPU (one process per GPU, torch.distributed / NCCL): sequences are
sharded in contiguous blocks over ranks, parameters are replicated, and every
objective evaluation does ONE all-reduce of the gradient (+ one 8-byte
all-reduce of -loglk); the regulariser and the whole L-BFGS update then run
identically (bit-for-bit, deterministic reductions) on every rank.  The Hamming
pass shards the upper-triangular pair tiles and all-reduces the int32 counters.
"""
import ctypes
import os

import numpy as np

from . import _lib
from . import lbfgs as _lbfgs


# backward implementation of the data term: "gather" (shared-memory bucket kernel) or "tc" (tcgen05 GEMM)
DEFAULT_BACKWARD = "tc"
DEFAULT_FORWARD = "tc"


def _torch():
    import torch
    return torch


from .dist import Collective, shard_bounds   # noqa: E402,F401


class CudaEngine(object):
    def __init__(self, device=None, group=None):
        self.lib = _lib.load()
        _lib.require_device()
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.EngineUnavailableError("torch sees no CUDA device; the PLM engine has no CPU fallback")
        self.coll = Collective(group)
        self.rank, self.world = self.coll.rank, self.coll.world
        if device is None:
            device = torch.cuda.current_device()
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device)
        self.kernel_launches = 0

    # -- helpers ---------------------------------------------------------------------------
    def stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def all_reduce(self, tensor):
        self.coll.all_reduce_sum(tensor)

    @staticmethod
    def ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    # -- (b) Hamming reweighting -----------------------------------------------------------
    def hamming_counts(self, codes, min_identical):
        """codes: (N, L) uint8 numpy (replicated on every rank).  Returns int32 numpy counts."""
        torch = _torch()
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        N, L = codes.shape
        if int(codes.max(initial=0)) >= 32:
            raise ValueError("sequence codes must be < 32")
        d_codes = torch.from_numpy(codes).to(self.device)
        d_counts = self.hamming_counts_device(d_codes, N, L, min_identical)
        return d_counts.cpu().numpy()

    def hamming_counts_device(self, d_codes, N, L, min_identical):
        torch = _torch()
        lib = self.lib
        words = lib.evc_hamming_plane_words(N, L)
        d_planes = torch.empty(words, dtype=torch.int32, device=self.device)
        d_counts = torch.zeros(N, dtype=torch.int32, device=self.device)
        _lib.check(lib.evc_hamming_pack(self.ptr(d_codes), N, L, self.ptr(d_planes), self.stream()),
                   "evc_hamming_pack")
        ntiles = lib.evc_hamming_num_tiles(N)
        lo, hi = shard_bounds(ntiles, self.world, self.rank)
        _lib.check(lib.evc_hamming_count_tiles(self.ptr(d_planes), N, L, int(min_identical), lo, hi,
                                               self.ptr(d_counts), self.stream()), "evc_hamming_count_tiles")
        self.kernel_launches += 2
        self.all_reduce(d_counts)
        return d_counts

    # -- (a) PLM ---------------------------------------------------------------------------
    def plm_problem(self, codes, weights, q, gap_code, lambda_h, lambda_J, m=6, backward=None, forward=None):
        return CudaPlmProblem(self, codes, weights, q, gap_code, lambda_h, lambda_J, m, backward, forward)


class CudaPlmProblem(object):
    """PLM objective on this rank's sequence shard + the L-BFGS vector space
    (see lbfgs.py for the protocol).  All n-vectors are torch CUDA tensors."""

    def __init__(self, engine, codes, weights, q, gap_code, lambda_h, lambda_J, m=6, backward=None,
                 forward=None):
        torch = _torch()
        if forward is None:
            forward = os.environ.get("EVC_FORWARD", DEFAULT_FORWARD)
        if forward not in ("gather", "tc", "tcfused"):
            raise ValueError("forward must be 'gather', 'tc' or 'tcfused'")
        if forward in ("tc", "tcfused"):
            backward = "tc"
        self.forward = forward
        if backward is None:
            backward = os.environ.get("EVC_BACKWARD", DEFAULT_BACKWARD)
        if backward not in ("gather", "tc"):
            raise ValueError("backward must be 'gather' or 'tc'")
        self.backward = backward
        self.engine = engine
        self.lib = engine.lib
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        weights = np.ascontiguousarray(weights, dtype=np.float32)
        N, L = codes.shape
        if weights.shape != (N,):
            raise ValueError("weights must have one entry per sequence")
        self.N_total, self.L, self.q, self.gap_code = N, L, int(q), int(gap_code)
        self.lambda_h, self.lambda_J = float(lambda_h), float(lambda_J)
        lo, hi = shard_bounds(N, engine.world, engine.rank)
        if hi <= lo:
            raise ValueError("fewer sequences than ranks")
        self.shard = (lo, hi)
        c_shard = np.ascontiguousarray(codes[lo:hi])
        w_shard = np.ascontiguousarray(weights[lo:hi])
        self.handle = ctypes.c_void_p()
        _lib.check(self.lib.evc_plm_create(ctypes.byref(self.handle), c_shard.ctypes.data_as(ctypes.c_void_p),
                                           hi - lo, L, self.q, self.gap_code,
                                           w_shard.ctypes.data_as(ctypes.c_void_p), engine.device_index),
                   "evc_plm_create")
        if backward == "tc":
            _lib.check(self.lib.evc_plm_set_backward(self.handle, 1), "evc_plm_set_backward")
        if forward in ("tc", "tcfused"):
            _lib.check(self.lib.evc_plm_set_forward(self.handle, 2 if forward == "tcfused" else 1),
                       "evc_plm_set_forward")
        self.n = int(self.lib.evc_plm_num_params(self.handle))
        dev = engine.device
        self.m = m
        f32 = dict(dtype=torch.float32, device=dev)
        self.x = torch.zeros(self.n, **f32)
        self.g = torch.zeros(self.n, **f32)
        self.xp = torch.zeros(self.n, **f32)
        self.gp = torch.zeros(self.n, **f32)
        self.d = torch.zeros(self.n, **f32)
        self.S = torch.zeros((m, self.n), **f32)
        self.Y = torch.zeros((m, self.n), **f32)
        self.ys = torch.zeros(m, dtype=torch.float64, device=dev)
        self.scratch = torch.zeros(m + 2, dtype=torch.float64, device=dev)
        self.fxbuf = torch.zeros(2, dtype=torch.float64, device=dev)
        self.dotbuf = torch.zeros(1, dtype=torch.float64, device=dev)
        self.last_negloglk = float("nan")
        self.evaluations = 0
        # own kernels per evaluate(): expand, fwd, bwd, finalize pairs + fields, add_reg x2
        self.launches_per_eval = {"tc": 8, "tcfused": 7, "gather": 7}[forward]

    def close(self):
        if self.handle:
            self.lib.evc_plm_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- objective ----------------------------------------------------------------------------
    def evaluate_async(self, x):
        """Launch data term + all-reduce + regulariser on the current stream; results in self.g / self.fxbuf."""
        e, p = self.engine, self.engine.ptr
        _lib.check(self.lib.evc_plm_eval_data(self.handle, p(x), p(self.g), p(self.fxbuf), e.stream()),
                   "evc_plm_eval_data")
        if e.world > 1:
            e.all_reduce(self.g)
            e.all_reduce(self.fxbuf[0:1])
        _lib.check(self.lib.evc_plm_add_regulariser(self.handle, p(x), p(self.g), p(self.fxbuf),
                                                    self.lambda_h, self.lambda_J, e.stream()),
                   "evc_plm_add_regulariser")
        self.evaluations += 1
        e.kernel_launches += self.launches_per_eval

    def evaluate(self, x):
        self.evaluate_async(x)
        nll, fx = self.fxbuf.tolist()
        self.last_negloglk = nll
        return fx

    def evaluate_host(self, x_host, g_host):
        """Public host-buffer entry: x_host / g_host are CPU float32 torch tensors (pinned for speed).
        H2D of x, evaluation (incl. the all-reduce), D2H of the gradient and of fx, then synchronise."""
        self.x.copy_(x_host, non_blocking=True)
        self.evaluate_async(self.x)
        g_host.copy_(self.g, non_blocking=True)
        nll, fx = self.fxbuf.tolist()          # D2H of the result scalars; synchronises the stream
        self.last_negloglk = nll
        return fx

    # -- vector space protocol ------------------------------------------------------------------
    def dot(self, a, b):
        e = self.engine
        _lib.check(self.lib.evc_vec_dot(e.ptr(a), e.ptr(b), self.n, e.ptr(self.dotbuf), e.stream()), "evc_vec_dot")
        e.kernel_launches += 2
        return float(self.dotbuf.item())

    def copy(self, dst, src):
        e = self.engine
        _lib.check(self.lib.evc_vec_copy(e.ptr(dst), e.ptr(src), self.n, e.stream()), "evc_vec_copy")

    def axpby(self, y, x, a, b):
        e = self.engine
        _lib.check(self.lib.evc_vec_axpby(e.ptr(y), e.ptr(x), float(a), float(b), self.n, e.stream()),
                   "evc_vec_axpby")
        e.kernel_launches += 1

    def update_pair(self, slot, xp, gp):
        e, p = self.engine, self.engine.ptr
        _lib.check(self.lib.evc_lbfgs_update_pair(p(self.S[slot]), p(self.Y[slot]), p(self.x), p(xp), p(self.g),
                                                  p(gp), p(self.ys[slot:slot + 1]), p(self.scratch[0:1]),
                                                  self.n, e.stream()), "evc_lbfgs_update_pair")
        e.kernel_launches += 3

    def direction(self, d, bound, end):
        e, p = self.engine, self.engine.ptr
        _lib.check(self.lib.evc_lbfgs_direction(p(d), p(self.g), p(self.S), p(self.Y), p(self.ys), p(self.scratch),
                                                self.n, self.m, int(bound), int(end), e.stream()),
                   "evc_lbfgs_direction")
        e.kernel_launches += 2 + 6 * int(bound)

    # -- a6 / a10 ---------------------------------------------------------------------------------
    def weighted_counts(self):
        """Returns (fi_counts (L,q), fij_counts (npairs,q,q)) float64 numpy, summed over all ranks."""
        torch = _torch()
        e = self.engine
        L, q = self.L, self.q
        buf = torch.zeros(self.n, dtype=torch.float32, device=e.device)
        _lib.check(self.lib.evc_plm_weighted_counts(self.handle, e.ptr(buf), e.ptr(buf[L * q:]), e.stream()),
                   "evc_plm_weighted_counts")
        e.kernel_launches += 5
        e.all_reduce(buf)
        host = buf.cpu().numpy().astype(np.float64)
        return host[:L * q].reshape(L, q), host[L * q:].reshape(L * (L - 1) // 2, q, q)

    def fn_scores(self, x=None):
        torch = _torch()
        e = self.engine
        x = self.x if x is None else x
        L, q = self.L, self.q
        out = torch.zeros(L * (L - 1) // 2, dtype=torch.float32, device=e.device)
        _lib.check(self.lib.evc_fn_scores(e.ptr(x[L * q:]), L, q, e.ptr(out), e.stream()), "evc_fn_scores")
        e.kernel_launches += 1
        return out.cpu().numpy()

    def set_x(self, x_host):
        torch = _torch()
        self.x.copy_(torch.from_numpy(np.ascontiguousarray(x_host, dtype=np.float32)))

    def get_x(self):
        return self.x.cpu().numpy()

    def norms(self):
        """(|h|, |J|) of the current parameters (for the iteration table), via the library's dot kernels."""
        import math
        e = self.engine
        nh = self.L * self.q
        out = []
        for lo, n in ((0, nh), (nh, self.n - nh)):
            v = self.x[lo:lo + n]
            _lib.check(self.lib.evc_vec_dot(e.ptr(v), e.ptr(v), n, e.ptr(self.dotbuf), e.stream()), "evc_vec_dot")
            e.kernel_launches += 2
            out.append(math.sqrt(float(self.dotbuf.item())))
        return out[0], out[1]

    def fit(self, x0, params, progress=None):
        self.set_x(x0)
        return _lbfgs.minimize(self, params, progress)
