"""
Generate the committed golden fixtures in tests/golden/.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

Sources of truth
  (1) the real plmc run shipped with the reference:
      /root/reference/notebooks/example/PABP_YEAST.{a2m,model_params}, PABP_YEAST_ECs.txt
      (presumed command: plmc -f PABP_YEAST -g -m 200 -t 0.2 -lh 0.01 -le 16.2)
  (2) the reference's own Python, imported unmodified via ref_harness:
      evcouplings/align/alignment.py:1192-1233 num_cluster_members,
      :1078-1153 frequencies / pair_frequencies,
      evcouplings/couplings/model.py:317-400 CouplingsModel reader (+ cn/fn scores :744-827),
      evcouplings/couplings/tools.py:20-108 parse_plmc_log.
Nothing from the reference's *source code* is copied; only its outputs on
seeded inputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from oracle import plm_oracle as po  # noqa: E402

EX = "/root/reference/notebooks/example"


def pabp():
    ids, seqs = po.read_a2m(os.path.join(EX, "PABP_YEAST.a2m"))
    prep = po.prepare_alignment(ids, seqs, focus="PABP_YEAST", alphabet=None, ignore_gaps=True)
    gm = po.read_model(os.path.join(EX, "PABP_YEAST.model_params"))
    counts_all = gm["weights"].astype(np.int32)           # golden stores integer neighbour counts
    np.savez_compressed(
        os.path.join(HERE, "pabp_codes.npz"),
        codes=prep["codes"], valid_packed=np.packbits(prep["valid"]),
        n_total=prep["n_total"], golden_counts_all=counts_all,
        focus_cols=prep["focus_cols"], index_list=prep["index_list"],
        target_seq=np.array(prep["target_seq"]), region_start=prep["region_start"],
    )
    # EC text
    ec = np.loadtxt(os.path.join(EX, "PABP_YEAST_ECs.txt"), dtype=str)
    pairs = [(0, 1), (6, 8), (3, 60), (20, 21), (33, 70), (50, 81)]
    L = gm["L"]
    iu, ju = np.triu_indices(L, 1)
    pidx = np.array([np.nonzero((iu == i) & (ju == j))[0][0] for i, j in pairs])
    # reference reader KATs on the golden file
    ref_harness.install()
    from evcouplings.couplings.model import CouplingsModel
    cm = CouplingsModel(os.path.join(EX, "PABP_YEAST.model_params"))
    kat = dict(
        hi_127=float(cm.hi(127, cm.seq(127))),
        Jij_127_172=float(cm.Jij(127, 172, cm.seq(127), cm.seq(172))),
        ref_cn_zero_sum=cm.cn_scores[iu, ju].astype(np.float64),   # model.py:788-803 (zero-sum gauge first)
    )
    np.savez_compressed(
        os.path.join(HERE, "pabp_golden.npz"),
        hdr_i=np.array([gm["L"], gm["q"], gm["n_valid"], gm["n_invalid"], gm["num_iter"]], dtype=np.int32),
        hdr_f=np.array([gm["theta"], gm["lambda_h"], gm["lambda_J"], gm["lambda_group"], gm["n_eff"]],
                       dtype=np.float32),
        alphabet=np.array(gm["alphabet"]), target_seq=np.array(gm["target_seq"]),
        index_list=gm["index_list"], fi=gm["fi"], h=gm["h"], J=gm["J"],
        fij_pairs=np.array(pairs, dtype=np.int32), fij_pair_index=pidx, fij_blocks=gm["fij"][pidx],
        ec_i=ec[:, 0].astype(np.int32), ec_Ai=ec[:, 1], ec_j=ec[:, 2].astype(np.int32), ec_Aj=ec[:, 3],
        ec_cn=ec[:, 5].astype(np.float64),
        kat_hi_127=kat["hi_127"], kat_Jij_127_172=kat["Jij_127_172"],
        ref_cn_zero_sum=kat["ref_cn_zero_sum"],
    )
    print("pabp fixtures written; valid=%d" % prep["n_valid"])


def in_tree_twins():
    """Reference numba kernels on the seeded config-1 alignment (N=200, L=40, q=21)."""
    ref_harness.install()
    from evcouplings.align import alignment as al
    out = {}
    for name, (N, L, seed, theta) in dict(cfg1=(200, 40, 1, 0.8), tie=(300, 50, 7, 0.8),
                                          odd=(257, 33, 11, 0.7)).items():
        codes = po.synthetic_msa_codes(N, L, seed)
        m = codes.astype(np.int64)
        counts = al.num_cluster_members(m, theta)
        w = 1.0 / counts
        fi = al.frequencies(m, w, 21)
        fij = al.pair_frequencies(m, w, 21, fi)
        iu, ju = np.triu_indices(L, 1)
        out[name + "_codes"] = codes
        out[name + "_theta"] = theta
        out[name + "_counts"] = counts.astype(np.int32)
        out[name + "_fi"] = fi
        out[name + "_fij_tri"] = fij[iu, ju]
    np.savez_compressed(os.path.join(HERE, "intree_twins.npz"), **out)
    print("in-tree twin fixtures written")


def tiny_model():
    """Tiny model written in plmc_v2 layout, read back by the reference's CouplingsModel."""
    ref_harness.install()
    from evcouplings.couplings.model import CouplingsModel
    N, L, q = 60, 12, 21
    codes = po.synthetic_msa_codes(N, L, 21)
    counts = po.hamming_counts(codes, 0.8)
    w = po.sequence_weights(counts)
    fi, fij = po.frequencies(codes, w, q)
    x, res = po.fit(codes, w, q, 0.01, 0.01 * (q - 1) * (L - 1), max_iter=400)
    h, Jt = po.unpack(x, L, q)
    alphabet = po.ALPHABET_PROTEIN
    target = "".join(alphabet[c] for c in codes[0])
    index_list = np.arange(5, 5 + L, dtype=np.int32)
    path = os.path.join(HERE, "tiny.model")
    po.write_model(path, L, q, N, 0, int(res.nit), 0.2, 0.01, 0.01 * (q - 1) * (L - 1), 0.0, w.sum(),
                   alphabet, counts.astype(np.float32), target, index_list, fi, h, fij, Jt)
    ecs_path = os.path.join(HERE, "tiny_ECs.txt")
    po.write_ecs(ecs_path, Jt.astype(np.float32), L, index_list, target)
    cm = CouplingsModel(path)
    iu, ju = np.triu_indices(L, 1)
    from evcouplings.couplings.pairs import read_raw_ec_file
    ecs = read_raw_ec_file(ecs_path, sort=False)
    np.savez_compressed(
        os.path.join(HERE, "tiny_ref_read.npz"),
        codes=codes, counts=counts, x=x.astype(np.float64),
        ref_J_tri=cm.J_ij[iu, ju], ref_h=cm.h_i, ref_fi=cm.f_i, ref_fij_tri=cm.f_ij[iu, ju],
        ref_L=cm.L, ref_q=cm.num_symbols, ref_N_eff=cm.N_eff, ref_theta=cm.theta,
        ref_lambda_J=cm.lambda_J, ref_alphabet=np.array("".join(cm.alphabet)),
        ref_target=np.array("".join(cm.target_seq)), ref_index_list=cm.index_list,
        ref_fn=cm.fn_scores[iu, ju], ref_cn_zero_sum=cm.cn_scores[iu, ju],
        ecs_i=ecs["i"].values, ecs_j=ecs["j"].values, ecs_cn=ecs["cn"].values,
        ecs_Ai=ecs["A_i"].values.astype(str), ecs_Aj=ecs["A_j"].values.astype(str),
    )
    print("tiny model fixtures written; iters=%d" % res.nit)


def model_consumers():
    """Reference CouplingsModel outputs (model.py) on the tiny model and on the golden PABP model:
    ecs table scores, hamiltonians, single-mutant matrix, delta_hamiltonian."""
    ref_harness.install()
    from evcouplings.couplings.model import CouplingsModel
    rng = np.random.default_rng(5)
    out = {}
    for name, path in (("tiny", os.path.join(HERE, "tiny.model")),
                       ("pabp", os.path.join(EX, "PABP_YEAST.model_params"))):
        cm = CouplingsModel(path)
        L, q = cm.L, cm.num_symbols
        iu, ju = np.triu_indices(L, 1)
        alphabet = "".join(cm.alphabet)
        tgt = "".join(cm.target_seq)
        seqs = [tgt] + ["".join(alphabet[k] for k in rng.integers(0, q, L)) for _ in range(40)]
        # a few sequences close to the target
        for _ in range(20):
            s = list(tgt)
            for p in rng.integers(0, L, 3):
                s[p] = alphabet[rng.integers(0, q)]
            seqs.append("".join(s))
        out[name + "_seqs"] = np.array(seqs)
        out[name + "_H"] = cm.hamiltonians(seqs)
        out[name + "_smm"] = cm.single_mut_mat_full
        out[name + "_fn"] = cm.fn_scores[iu, ju]
        out[name + "_cn"] = cm.cn_scores[iu, ju]
        if name == "tiny":
            out[name + "_mi_raw"] = cm.mi_scores_raw[iu, ju]
            out[name + "_mi_apc"] = cm.mi_scores_apc[iu, ju]
        variants = []
        for _ in range(12):
            ps = sorted(set(int(p) for p in rng.integers(0, L, 2)))
            variants.append([(int(cm.index_list[p]), tgt[p], alphabet[rng.integers(0, q)]) for p in ps])
        out[name + "_var_pos"] = np.array([[v[0][0], v[-1][0]] for v in variants])
        out[name + "_variants"] = np.array([";".join("%d,%s,%s" % s for s in v) for v in variants])
        out[name + "_dH"] = np.array([cm.delta_hamiltonian(v) for v in variants])
    np.savez_compressed(os.path.join(HERE, "model_consumers.npz"), **out)
    print("model consumer fixtures written; PABP H(target) = %.10f, smm(127,E) = %.10f" % (
        out["pabp_H"][0, 0], out["pabp_smm"][127 - 123, "ACDEFGHIKLMNPQRSTVWY".index("E"), 0]))


if __name__ == "__main__":
    pabp()
    in_tree_twins()
    tiny_model()
    model_consumers()
