"""
GPU drop-ins for the reference's in-tree alignment statistics: the single-threaded
numba twins of hot path (b) and of the frequency counting that the reference runs in its align stage
(``Alignment.set_weights`` evcouplings/align/alignment.py:899-930, called from align/protocol.py:966) and in
mean-field DCA (couplings/mean_field.py:187-195).  Same argument meaning and return values as the
reference functions they replace; the work runs in the kernels of libevcplm (no CPU fallback).

    num_cluster_members(matrix, identity_threshold)   <- alignment.py:1192-1233
    frequencies(matrix, seq_weights, num_symbols)     <- alignment.py:1078-1106
    pair_frequencies(matrix, seq_weights, num_symbols, fi)  <- alignment.py:1109-1153
    identities_to_seq(seq, matrix)                    <- alignment.py:1156-1189
    set_weights(alignment, identity_threshold)        <- Alignment.set_weights :899-930 (duck-typed)
"""
import numpy as np

from . import msa

_engine = None


def _get_engine():
    global _engine
    if _engine is None:
        from .engine import CudaEngine
        _engine = CudaEngine()
    return _engine


def num_cluster_members(matrix, identity_threshold, engine=None):
    """Number of sequences within ``identity_threshold`` of each sequence (self included); ``matrix`` is the
    N x L integer matrix produced by the reference's ``map_matrix``.  Returns float64 like the reference."""
    eng = engine or _get_engine()
    m = np.ascontiguousarray(matrix)
    if m.ndim != 2:
        raise ValueError("matrix must be N x L")
    if m.size and (m.min() < 0 or m.max() >= 32):
        raise ValueError("mapped symbols must be in [0, 32)")
    thr = msa.identity_threshold_count(float(identity_threshold), m.shape[1])
    return eng.hamming_counts(m.astype(np.uint8), thr).astype(np.float64)


def _as_codes(matrix, num_symbols):
    """mapped integer matrix -> uint8 codes, range-checked BEFORE the narrowing cast (a symbol >= num_symbols
    would index another site's coupling block inside the kernels)"""
    m = np.ascontiguousarray(matrix)
    if m.ndim != 2:
        raise ValueError("matrix must be N x L")
    if m.size and (m.min() < 0 or m.max() >= int(num_symbols)):
        raise ValueError("mapped symbols must be in [0, %d)" % int(num_symbols))
    return m.astype(np.uint8)


def identities_to_seq(seq, matrix, engine=None):
    """Number of identities of every sequence in ``matrix`` (N x L, mapped) to ``seq`` (length L, mapped).
    Returns float64 of length N like the reference twin (alignment.py:1156-1189)."""
    import ctypes
    import torch
    from . import _lib
    eng = engine or _get_engine()
    m = _as_codes(matrix, 256)
    s = np.ascontiguousarray(seq)
    N, L = m.shape
    if s.shape != (L,):
        raise ValueError("seq must have one entry per column of matrix")
    if N == 0:
        return np.zeros(0)
    if s.min() < 0 or s.max() > 255:
        raise ValueError("mapped symbols must be in [0, 256)")
    d_m = torch.from_numpy(m).to(eng.device)
    d_s = torch.from_numpy(s.astype(np.uint8)).to(eng.device)
    out = torch.zeros(N, dtype=torch.int32, device=eng.device)
    _lib.check(eng.lib.evc_identities_to_seq(eng.ptr(d_m), eng.ptr(d_s), N, L, eng.ptr(out), eng.stream()),
               "evc_identities_to_seq")
    eng.kernel_launches += 1
    return out.cpu().numpy().astype(np.float64)


def _problem(matrix, seq_weights, num_symbols, engine):
    eng = engine or _get_engine()
    m = _as_codes(matrix, num_symbols)
    w = np.ascontiguousarray(seq_weights, dtype=np.float32)
    return eng.plm_problem(m, w, int(num_symbols), -1, 0.0, 0.0, forward="gather", backward="gather")


def frequencies(matrix, seq_weights, num_symbols, engine=None):
    """Single-site frequencies, L x num_symbols, normalised by the sum of weights (alignment.py:1106)."""
    prob = _problem(matrix, seq_weights, num_symbols, engine)
    try:
        fi, _ = prob.weighted_counts()
    finally:
        prob.close()
    return fi / float(np.sum(np.asarray(seq_weights, dtype=np.float64)))


def pair_frequencies(matrix, seq_weights, num_symbols, fi, engine=None):
    """Pair frequencies, L x L x q x q, with the reference's conventions: symmetric fill,
    f_ij[i, i, a, a] = f_i[i, a] on the diagonal (alignment.py:1144-1151)."""
    prob = _problem(matrix, seq_weights, num_symbols, engine)
    try:
        _, fij_tri = prob.weighted_counts()
    finally:
        prob.close()
    L = np.asarray(matrix).shape[1]
    q = int(num_symbols)
    neff = float(np.sum(np.asarray(seq_weights, dtype=np.float64)))
    fij = np.zeros((L, L, q, q))
    iu, ju = np.triu_indices(L, 1)
    fij[iu, ju] = fij_tri / neff
    fij[ju, iu] = (fij_tri / neff).transpose(0, 2, 1)
    idx = np.arange(q)
    for i in range(L):
        fij[i, i, idx, idx] = np.asarray(fi)[i]
    return fij


def set_weights(alignment, identity_threshold=0.8, engine=None):
    """Drop-in for ``Alignment.set_weights``: fills ``alignment.num_cluster_members`` and ``alignment.weights``
    (an object with ``matrix_mapped`` / ``__ensure_mapped_matrix`` semantics of the reference's Alignment)."""
    ensure = getattr(alignment, "_Alignment__ensure_mapped_matrix", None)      # the reference's private helper
    if ensure is not None:
        ensure()
    mapped = getattr(alignment, "matrix_mapped", None)
    if mapped is None:
        from evcouplings.align.alignment import map_matrix
        mapped = map_matrix(alignment.matrix, alignment.alphabet_map)
        alignment.matrix_mapped = mapped
    alignment.num_cluster_members = num_cluster_members(mapped, identity_threshold, engine)
    alignment.weights = 1.0 / alignment.num_cluster_members
    # like the reference (alignment.py:926-930): cached frequencies were computed with other / no weights
    alignment._frequencies = None
    alignment._pair_frequencies = None
    return alignment
