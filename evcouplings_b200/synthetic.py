"""
Deterministic synthetic alignments of the BASELINE shapes: K = ceil(N/50) random
family centres over the 20 residues, each sequence a copy of a random centre with per-sequence
mutation probability p ~ U(0.1, 0.6) and per-site gap probability 0.05; row 0 (the focus) is gap-free.
Codes are in gap-as-state convention (0 = gap, 1..20 = ACDEFGHIKLMNPQRSTVWY).
"""
import numpy as np

ALPHABET = "-ACDEFGHIKLMNPQRSTVWY"

CONFIG_SEEDS = {1: (200, 40, 1), 2: (50000, 200, 2), 3: (200000, 300, 3), 4: (500000, 500, 4),
                5: (100000, 800, 5)}


def synthetic_msa_codes(N, L, seed, q_res=20, gap_prob=0.05):
    rng = np.random.default_rng(seed)
    K = max(1, -(-N // 50))
    centres = rng.integers(1, q_res + 1, size=(K, L), dtype=np.uint8)
    which = rng.integers(0, K, size=N)
    p_mut = rng.uniform(0.1, 0.6, size=N)
    codes = centres[which]
    mut = rng.random((N, L)) < p_mut[:, None]
    rnd = rng.integers(1, q_res + 1, size=(N, L), dtype=np.uint8)
    codes = np.where(mut, rnd, codes)
    gaps = rng.random((N, L)) < gap_prob
    gaps[0, :] = False
    return np.where(gaps, 0, codes).astype(np.uint8)


def to_ignore_gaps_codes(codes, q=20):
    """gap-as-state codes (gap 0, residues 1..20) -> ignore_gaps codes (residues 0..19, gap 20)."""
    return np.where(codes == 0, q, codes - 1).astype(np.uint8)


def write_a2m(path, codes, focus_name="seq0"):
    N, L = codes.shape
    lut = np.frombuffer(ALPHABET.encode("ascii"), dtype=np.uint8)
    chars = lut[codes]
    with open(path, "w") as f:
        for n in range(N):
            name = "%s/1-%d" % (focus_name, L) if n == 0 else "seq%d/1-%d" % (n, L)
            f.write(">%s\n%s\n" % (name, bytes(chars[n]).decode("ascii")))
