"""f4 timing: compiled A2M ingest (csrc/a2m_reader.cu) vs the pure-Python line loop at Pfam scale (N=500k, L=500).
Host-only; writes profiles/r2_ingest_timing.json."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evcouplings_b200 import msa, synthetic  # noqa: E402

N, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (500000, 500)
codes = synthetic.synthetic_msa_codes(N, L, 4)
d = tempfile.mkdtemp()
path = os.path.join(d, "cfg4.a2m")
synthetic.write_a2m(path, codes)
size = os.path.getsize(path)
out = {"N": N, "L": L, "file_bytes": size, "host_threads": len(os.sched_getaffinity(0))}
for rep in range(2):
    t0 = time.perf_counter(); ids, raw = msa.read_fasta_matrix(path); t1 = time.perf_counter()
    ali = msa.encode_alignment(ids, raw, focus="seq0"); t2 = time.perf_counter()
    out["compiled"] = {"read_s": t1 - t0, "encode_s": t2 - t1, "total_s": t2 - t0, "GBps": size / (t2 - t0) / 1e9}
assert np.array_equal(ali.codes, codes) and ali.n_valid == N
t0 = time.perf_counter(); ids2, raw2 = msa.read_fasta_matrix_py(path); t1 = time.perf_counter()
assert ids2 == ids and np.array_equal(raw2, raw)
out["python_line_loop"] = {"read_s": t1 - t0}
os.unlink(path); os.rmdir(d)
print(json.dumps(out))
with open(os.path.join(ROOT, "profiles", "r2_ingest_timing.json"), "w") as f:
    json.dump(out, f, indent=1)
