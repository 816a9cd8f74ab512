"""
SURVEY.md 8(f) rows f1 / f2: GPU versions of what the reference's ``CouplingsModel`` does with a fitted
model (evcouplings/couplings/model.py) -- same numbers, same table layout, no per-pair Python loops:

* ``read_model``            bulk plmc_v2 reader (the reference issues L(L-1) np.fromfile calls, model.py:375-389)
* ``ec_table``              FN / CN (zero-sum gauge + APC) / MI scores  <- _calculate_ecs  model.py:777-827
* ``hamiltonians``          statistical energies of many sequences     <- _hamiltonians    model.py:25-60
* ``single_mutant_matrix``  all single substitutions of the target     <- _single_mutant_hamiltonians model.py:63-109
* ``delta_hamiltonians``    energies of variants relative to the target <- delta_hamiltonian model.py:672-712
"""
import ctypes

import numpy as np

from . import _lib


def read_model(path):
    """Bulk reader of the plmc_v2 layout (model.py:317-389); tri blocks stay packed (npairs, q, q)."""
    with open(path, "rb") as f:
        L, q, nv, ni, it = (int(v) for v in np.fromfile(f, "<i4", 5))
        theta, lh, lj, lg, neff = (float(v) for v in np.fromfile(f, "<f4", 5))
        alphabet = f.read(q).decode("ascii")
        weights = np.fromfile(f, "<f4", nv + ni)
        target = f.read(L).decode("ascii")
        index_list = np.fromfile(f, "<i4", L)
        fi = np.fromfile(f, "<f4", L * q).reshape(L, q)
        h = np.fromfile(f, "<f4", L * q).reshape(L, q)
        npair = L * (L - 1) // 2
        fij = np.fromfile(f, "<f4", npair * q * q).reshape(npair, q, q)
        J = np.fromfile(f, "<f4", npair * q * q).reshape(npair, q, q)
    if J.size != npair * q * q:
        raise ValueError("truncated model file: " + str(path))
    return dict(L=L, q=q, n_valid=nv, n_invalid=ni, num_iter=it, theta=theta, lambda_h=lh, lambda_J=lj,
                lambda_group=lg, n_eff=neff, alphabet=alphabet, weights=weights, target_seq=target,
                index_list=index_list, fi=fi, h=h, fij=fij, J=J)


def _engine(engine):
    if engine is not None:
        return engine
    from .engine import CudaEngine
    return CudaEngine()


def apc(matrix):
    """Average product correction exactly as model.py:744-775 (diagonal blanked)."""
    L = matrix.shape[0]
    col_means = np.mean(matrix, axis=0) * L / (L - 1)
    matrix_mean = np.mean(matrix) * L / (L - 1)
    out = matrix - np.outer(col_means, col_means) / matrix_mean
    out[np.diag_indices(L)] = 0
    return out


def pair_scores(model, engine=None):
    """Per-pair raw-gauge FN, zero-sum-gauge FN and MI (pair order i<j row-major) from the device."""
    import torch
    eng = _engine(engine)
    L, q = model["L"], model["q"]
    dev = eng.device
    J = torch.from_numpy(np.ascontiguousarray(model["J"], dtype=np.float32)).to(dev)
    fij = torch.from_numpy(np.ascontiguousarray(model["fij"], dtype=np.float32)).to(dev)
    fi = torch.from_numpy(np.ascontiguousarray(model["fi"], dtype=np.float32)).to(dev)
    npair = L * (L - 1) // 2
    out = torch.zeros((3, npair), dtype=torch.float32, device=dev)
    _lib.check(eng.lib.evc_ec_scores(eng.ptr(J), eng.ptr(fij), eng.ptr(fi), L, q, eng.ptr(out[0]), eng.ptr(out[1]),
                                     eng.ptr(out[2]), eng.stream()), "evc_ec_scores")
    eng.kernel_launches += 1
    o = out.cpu().numpy().astype(np.float64)
    return o[0], o[1], o[2]


def ec_table(model, engine=None):
    """DataFrame with the columns of CouplingsModel.ecs (model.py:806-827), sorted by cn descending."""
    import pandas as pd
    L = model["L"]
    fn_raw, fn_zs, mi = pair_scores(model, engine)
    iu, ju = np.triu_indices(L, 1)

    def full(v):
        m = np.zeros((L, L))
        m[iu, ju] = v
        return m + m.T

    cn = apc(full(fn_zs))[iu, ju]
    mi_apc = apc(full(mi))[iu, ju]
    idx = np.asarray(model["index_list"])
    tgt = model["target_seq"]
    df = pd.DataFrame({
        "i": idx[iu], "A_i": [tgt[k] for k in iu], "j": idx[ju], "A_j": [tgt[k] for k in ju],
        "seqdist": np.abs(idx[iu] - idx[ju]), "mi_raw": mi, "mi_apc": mi_apc, "fn": fn_zs, "cn": cn,
    })
    return df.sort_values(by="cn", ascending=False)


def encode_sequences(model, sequences):
    """list of strings -> (N, L) uint8 codes in the model alphabet; characters outside it become q (ignored)."""
    q = model["q"]
    lut = np.full(256, q, dtype=np.uint8)
    for k, ch in enumerate(model["alphabet"]):
        lut[ord(ch)] = k
    arr = np.frombuffer("".join(sequences).encode("ascii"), dtype=np.uint8).reshape(len(sequences), model["L"])
    return lut[arr]


def hamiltonians(model, sequences, engine=None):
    """(N, 3) float64: total, couplings and fields part of the statistical energy of every sequence
    (strings, or an (N, L) integer matrix already mapped to the model alphabet)."""
    import torch
    eng = _engine(engine)
    if len(sequences) and isinstance(sequences[0], str):
        codes = encode_sequences(model, sequences)
    else:
        arr = np.ascontiguousarray(sequences)
        if arr.size and (arr.min() < 0 or arr.max() > model["q"]):
            raise ValueError("mapped sequence symbols must be in [0, %d] (%d = gap)" % (model["q"], model["q"]))
        codes = arr.astype(np.uint8)
    N, L = codes.shape
    q = model["q"]
    gap_code = q if int(codes.max(initial=0)) >= q else -1
    w = np.ones(N, dtype=np.float32)
    handle = ctypes.c_void_p()
    _lib.check(eng.lib.evc_plm_create(ctypes.byref(handle), codes.ctypes.data_as(ctypes.c_void_p), N, L, q, gap_code,
                                      w.ctypes.data_as(ctypes.c_void_p), eng.device_index), "evc_plm_create")
    try:
        x = np.concatenate([np.asarray(model["h"], dtype=np.float32).ravel(),
                            np.asarray(model["J"], dtype=np.float32).ravel()])
        dx = torch.from_numpy(x).to(eng.device)
        out = torch.zeros((N, 3), dtype=torch.float64, device=eng.device)
        _lib.check(eng.lib.evc_plm_energies(handle, eng.ptr(dx), eng.ptr(out), eng.stream()), "evc_plm_energies")
        eng.kernel_launches += 3
        res = out.cpu().numpy()
    finally:
        eng.lib.evc_plm_destroy(handle)
    return res


def single_mutant_matrix(model, engine=None):
    """(L, q, 3) energy differences of every single substitution of the target sequence
    (model.py:63-109), obtained as energies of the L*q single mutants minus the target's."""
    L, q = model["L"], model["q"]
    tgt = encode_sequences(model, [model["target_seq"]])[0]
    muts = np.repeat(tgt[None, :], L * q + 1, axis=0)
    for i in range(L):
        muts[1 + i * q: 1 + (i + 1) * q, i] = np.arange(q)
    H = hamiltonians(model, muts, engine)
    return (H[1:] - H[0]).reshape(L, q, 3)


def delta_hamiltonians(model, variants, engine=None):
    """variants: list of substitution lists [(pos, from, to), ...] in index_list numbering (model.py:672-712).
    Returns (len(variants), 3)."""
    pos_of = {int(p): k for k, p in enumerate(model["index_list"])}
    amap = {ch: k for k, ch in enumerate(model["alphabet"])}
    tgt = encode_sequences(model, [model["target_seq"]])[0]
    seqs = np.repeat(tgt[None, :], len(variants) + 1, axis=0)
    for v, subs in enumerate(variants):
        for (p, a_from, a_to) in subs:
            k = pos_of[int(p)]
            if a_from != model["target_seq"][k]:
                raise ValueError("Inconsistency with target sequence: pos={} target={} subs={}".format(
                    p, model["target_seq"][k], a_from))
            seqs[v + 1, k] = amap[a_to]
    H = hamiltonians(model, seqs, engine)
    return H[1:] - H[0]
