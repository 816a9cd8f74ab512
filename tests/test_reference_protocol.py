"""
Drop-in boundary test (CPU, this container only): the reference's OWN couplings protocol
(evcouplings/couplings/protocol.py:363-429 ``standard`` -> ``infer_plmc`` :56-257) runs unmodified with
``evcouplings.couplings.tools.run_plmc`` replaced by ``evcouplings_b200.run_plmc``; the reference's own
readers (CouplingsModel model.py:317-400, read_raw_ec_file pairs.py:34-65, parse_plmc_log tools.py:20-108)
consume what we write.  The numerical engine injected here is the test-only oracle engine (no GPU in this
container); the same host code runs over the CUDA engine in tests/test_gpu_parity.py.
Skipped where /root/reference does not exist (the GPU box).
"""
import functools
import os

import numpy as np
import pytest

import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference not present (/root/reference or baseline/_ref)")


@pytest.fixture(scope="module")
def ref():
    ref_harness.install()
    import evcouplings.couplings.tools as ct
    import evcouplings.couplings.protocol as cpr
    import evcouplings.couplings.model as cm
    import evcouplings.couplings.pairs as cp
    return dict(ct=ct, cpr=cpr, cm=cm, cp=cp)


def _kwargs(prefix, a2m, L, ignore_gaps):
    return dict(
        protocol="standard", prefix=prefix, alignment_file=a2m, focus_mode=True, focus_sequence="seq0/1-%d" % L,
        theta=0.8, alphabet=None, segments=[["A_1", "aa", "seq0", 1, L, list(range(1, L + 1))]],
        ignore_gaps=ignore_gaps, iterations=30, lambda_h=0.01, lambda_J=0.01, lambda_J_times_Lq=True,
        lambda_group=None, scale_clusters=None, cpu=2, plmc="plmc", reuse_ecs=False, min_sequence_distance=6,
        frequencies_file=None, scoring_model="skewnormal",
    )


@pytest.mark.parametrize("ignore_gaps", [True, False])
def test_reference_standard_protocol_over_our_run_plmc(ref, tmp_path, ignore_gaps):
    from evcouplings_b200 import synthetic, tools
    from cpu_engine import OracleEngine
    from oracle import plm_oracle as po
    N, L = 200, 40                      # BASELINE configs[0]
    codes = synthetic.synthetic_msa_codes(N, L, 1)
    a2m = str(tmp_path / "cfg1.a2m")
    synthetic.write_a2m(a2m, codes)
    captured = {}

    def run_plmc(*args, **kwargs):
        res, run = tools.run_plmc(*args, engine=OracleEngine(), return_run=True, **kwargs)
        captured["run"], captured["kwargs"], captured["args"] = run, kwargs, args
        return res

    ct = ref["ct"]
    original = ct.run_plmc
    ct.run_plmc = run_plmc
    try:
        prefix = str(tmp_path / "out" / "job")
        outcfg = ref["cpr"].run(**_kwargs(prefix, a2m, L, ignore_gaps))
    finally:
        ct.run_plmc = original

    # the protocol handed us lambda_J already scaled by (q_eff - 1) * (L - 1)   (protocol.py:157-179)
    q_eff = 20 if ignore_gaps else 21
    assert abs(captured["kwargs"]["lambda_J"] - 0.01 * (q_eff - 1) * (L - 1)) < 1e-12
    assert captured["kwargs"]["focus_seq"] == "seq0/1-40" and captured["kwargs"]["theta"] == 0.8

    # stage outputs the rest of the pipeline consumes
    for key in ("model_file", "raw_ec_file", "ec_file"):
        assert os.path.getsize(outcfg[key]) > 0
    assert outcfg["num_sites"] == L and outcfg["num_valid_sequences"] == N
    assert abs(outcfg["effective_sequences"] - captured["run"].n_eff) < 0.06
    assert outcfg["region_start"] == 1
    assert os.path.exists(prefix + "_iteration_table.csv")
    assert os.path.exists(prefix + ".couplings_standard_plmc.outcfg")     # restart record (YAML of PlmcResult)

    # the reference's own readers on our files
    model = ref["cm"].CouplingsModel(outcfg["model_file"])
    run = captured["run"]
    assert model.L == L and model.num_symbols == q_eff and model.N_valid == N
    assert "".join(model.alphabet) == ("ACDEFGHIKLMNPQRSTVWY" if ignore_gaps else "-ACDEFGHIKLMNPQRSTVWY")
    h = run.x[:L * q_eff].reshape(L, q_eff)
    assert np.allclose(model.h_i, h, atol=0, rtol=0)
    iu, ju = np.triu_indices(L, 1)
    J = run.x[L * q_eff:].reshape(-1, q_eff, q_eff)
    assert np.array_equal(model.J_ij[iu, ju], J.astype(np.float64))
    assert np.array_equal(model.J_ij[ju, iu], J.transpose(0, 2, 1).astype(np.float64))
    assert abs(model.theta - 0.2) < 1e-7 and abs(model.N_eff - run.n_eff) < 1e-2
    assert "".join(model.target_seq) == run.alignment.target_seq
    ecs = ref["cp"].read_raw_ec_file(outcfg["raw_ec_file"], sort=False)
    assert len(ecs) == L * (L - 1) // 2
    assert np.abs(ecs["cn"].values - po.cn_scores(J, L)).max() < 1e-6
    # the reference's own log parser accepts our log and agrees with ours
    it_ref, fields_ref = ct.parse_plmc_log(run.log)
    it_own, fields_own = tools.parse_plmc_log(run.log)
    assert fields_ref == fields_own
    assert list(it_ref.columns) == list(it_own.columns) and len(it_ref) == len(it_own) == 30
    assert it_ref.equals(it_own)


def test_reference_parse_of_realistic_failure_modes(ref):
    """mandatory log lines: the reference raises KeyError without them (tools.py:97-99); ours too."""
    from evcouplings_b200 import tools
    with pytest.raises(KeyError):
        ref["ct"].parse_plmc_log("nothing useful")
    with pytest.raises(KeyError):
        tools.parse_plmc_log("nothing useful")


def test_product_ingest_on_real_pabp_alignment(golden_dir):
    """product ingest on the real A2M shipped with the reference == the golden fixture (which the oracle's
    per-character restatement produced and plmc's own header / weights confirm: 151,496 valid + 545 invalid)."""
    from evcouplings_b200 import msa
    path = os.path.join(ref_harness.REFERENCE_ROOT, "notebooks", "example", "PABP_YEAST.a2m")
    if not os.path.exists(path):
        pytest.skip("the example alignment ships only with the full reference checkout")
    ali = msa.load_alignment(path, focus="PABP_YEAST", ignore_gaps=True)
    c = np.load(os.path.join(golden_dir, "pabp_codes.npz"))
    valid = np.unpackbits(c["valid_packed"])[: int(c["n_total"])].astype(bool)
    assert np.array_equal(ali.codes, c["codes"]) and np.array_equal(ali.valid, valid)
    assert ali.target_seq == str(c["target_seq"]) and np.array_equal(ali.index_list, c["index_list"])
    assert (ali.n_valid, ali.n_total - ali.n_valid, ali.region_start, ali.num_total_sites) == (151496, 545, 115, 96)


def test_unmodified_reference_run_plmc_over_plmc_compatible_cli(ref, tmp_path):
    """Secondary plug point: the reference's OWN run_plmc (tools.py:126-307: argv, subprocess, stderr parsing,
    output checks) drives our plmc-compatible executable.  The wrapper used here injects the test-only oracle
    engine (no GPU in this container); bin/evcplm-plmc is the same entry point with the CUDA engine."""
    import stat
    import sys as _sys
    from evcouplings_b200 import synthetic
    from oracle import plm_oracle as po
    codes = synthetic.synthetic_msa_codes(150, 16, 3)
    a2m = str(tmp_path / "in.a2m")
    synthetic.write_a2m(a2m, codes)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    wrapper = tmp_path / "plmc_test_wrapper"
    wrapper.write_text(
        "#!%s\nimport sys\nsys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from cpu_engine import OracleEngine\nfrom evcouplings_b200.plmc_cli import main\n"
        "sys.exit(main(engine=OracleEngine()))\n" % (_sys.executable, root, os.path.join(root, "tests")))
    wrapper.chmod(wrapper.stat().st_mode | stat.S_IEXEC)
    ecs, model = str(tmp_path / "o" / "x_ECs.txt"), str(tmp_path / "o" / "x.model")
    res = ref["ct"].run_plmc(a2m, ecs, model, focus_seq="seq0/1-16", alphabet=None, theta=0.8, scale=None,
                             ignore_gaps=True, iterations=12, lambda_h=0.01, lambda_J=2.5, lambda_g=None, cpu=2,
                             binary=str(wrapper))
    assert res.num_valid_seqs == 150 and res.num_total_seqs == 150 and res.num_valid_sites == 16
    assert res.focus_seq_index == 1 and res.region_start == 1
    assert res.optimization_status == "LBFGSERR_MAXIMUMITERATION" and len(res.iteration_table) == 12
    m = po.read_model(model)
    assert (m["L"], m["q"], m["num_iter"]) == (16, 20, 12) and abs(m["theta"] - 0.2) < 1e-6
    assert abs(m["lambda_J"] - 2.5) < 1e-6 and abs(res.effective_samples - m["n_eff"]) < 0.06
    assert len(open(ecs).read().strip().split("\n")) == 16 * 15 // 2


def test_plmc_cli_argument_handling():
    from evcouplings_b200 import plmc_cli
    ali, o = plmc_cli.parse_args(["-c", "e.txt", "-o", "m.model", "-f", "SEQ", "-g", "-m", "100", "-t", "0.2",
                                  "-lh", "0.01", "-le", "16.2", "-n", "4", "in.a2m"])
    assert ali == "in.a2m" and o["couplings_file"] == "e.txt" and o["param_file"] == "m.model"
    assert o["focus_seq"] == "SEQ" and o["ignore_gaps"] and o["iterations"] == 100
    assert abs(o["theta"] - 0.8) < 1e-12 and o["lambda_h"] == 0.01 and o["lambda_J"] == 16.2 and o["cpu"] == "4"
    import io
    for bad in (["-c"], ["in.a2m"], ["-c", "e", "a", "b"], ["-zz", "1", "-c", "e", "a"]):
        assert plmc_cli.main(bad, stderr=io.StringIO()) == 2
    err = io.StringIO()
    assert plmc_cli.main(["-c", "/tmp/e.txt", "/nonexistent/file.a2m"], stderr=err) == 1 and "ResourceError" in err.getvalue()


def test_reference_complex_protocol_over_our_run_plmc(ref, tmp_path):
    """BASELINE configs[4] flavour (EVcomplex concatenated two-chain alignment): the reference's ``complex``
    protocol (protocol.py:480-594; same infer_plmc -> run_plmc boundary, two segments, inter-chain EC table)
    runs unmodified over our run_plmc.  Small shapes here (2 x 12 sites); the engine itself is parity- and
    bench-tested at L=800 on the GPU."""
    import pandas as pd
    from evcouplings_b200 import synthetic, tools
    from cpu_engine import OracleEngine
    N, L1, L2 = 160, 12, 12
    L = L1 + L2
    codes = synthetic.synthetic_msa_codes(N, L, 8)
    a2m = str(tmp_path / "complex.a2m")
    synthetic.write_a2m(a2m, codes, focus_name="A_B")          # header "A_B/1-24" like complex/alignment.py:85-92
    ct = ref["ct"]
    original = ct.run_plmc
    ct.run_plmc = lambda *a, **k: tools.run_plmc(*a, engine=OracleEngine(), **k)
    try:
        prefix = str(tmp_path / "cx" / "job")
        kw = _kwargs(prefix, a2m, L, True)
        kw.update(protocol="complex", focus_sequence="A_B/1-%d" % L, use_all_ecs_for_scoring=False,
                  segments=[["A_1", "aa", "A", 1, L1, list(range(1, L1 + 1))],
                            ["B_1", "aa", "B", 1, L2, list(range(1, L2 + 1))]])
        outcfg = ref["cpr"].run(**kw)
    finally:
        ct.run_plmc = original
    assert outcfg["num_sites"] == L and outcfg["num_valid_sequences"] == N
    inter = pd.read_csv(outcfg["inter_ec_file"])
    assert len(inter) == L1 * L2 and set(inter["segment_i"]) == {"A_1"} and set(inter["segment_j"]) == {"B_1"}
    allecs = pd.read_csv(outcfg["ec_file"])
    assert {"i", "j", "segment_i", "segment_j", "cn", "probability"} <= set(allecs.columns)
    model = ref["cm"].CouplingsModel(outcfg["model_file"])
    assert model.L == L and model.num_symbols == 20
