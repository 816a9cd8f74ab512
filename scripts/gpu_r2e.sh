#!/bin/bash
# Round-2 GPU batch E (1 GPU): regression after the per-(device, stream) scratch change; CTA-pair kernel with the lean
# control warps; line-search evaluations per iteration on the config-4 share; final default bench line.
cd "$(dirname "$0")/.."
O=gpurun_out/r2e; mkdir -p $O
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; rc=$?; tail -2 $O/smoke.txt
if [ $rc -ne 0 ]; then echo "SMOKE FAILED rc=$rc"; tail -30 $O/smoke.txt; exit 1; fi
echo "== full GPU test suite"
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_all.txt 2>&1; echo "rc=$?"; tail -5 $O/pytest_all.txt
echo "== CTA-pair kernel with lean control warps"
EVC_TC_PAIR=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/pair_smoke.txt 2>&1; rc=$?; tail -1 $O/pair_smoke.txt
if [ $rc -eq 0 ]; then
  EVC_TC_PAIR=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_core or large_L or full_size_properties or precision_bf16" > $O/pair_pytest.txt 2>&1; echo "rc=$?"; tail -2 $O/pair_pytest.txt
  for prec in fp32 bf16; do
    EVC_TC_PAIR=1 timeout 300 python bench.py --no-subrecords --steps 60 --precision $prec > $O/pair_bench_$prec.json 2>/dev/null
    python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})" $O/pair_bench_$prec.json
  done
fi
echo "== line-search behaviour on the config-4 share (62,500 x 500), 20 iterations"
timeout 900 python - > $O/fit_cfg4share.txt 2>&1 <<PY
import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np
from evcouplings_b200 import synthetic, tools
N, L = 62500, 500
codes = synthetic.synthetic_msa_codes(N, L, 4); a2m = "/tmp/cfg4s.a2m"; synthetic.write_a2m(a2m, codes)
for prec in ("fp32", "auto"):
    t0 = time.time()
    res, run = tools.run_plmc(a2m, "/tmp/cfg4s_ECs.txt", "/tmp/cfg4s.model", focus_seq="seq0", theta=0.8, iterations=20,
                              lambda_h=0.01, lambda_J=0.01 * 20 * (L - 1), num_gpus=1, return_run=True, precision=prec)
    tt = res.iteration_table["time"].astype(float).values
    print(prec, json.dumps(dict(wall_s=time.time() - t0, iterations=run.lbfgs.iterations, evaluations=run.lbfgs.evaluations,
          status=run.lbfgs.status, time_column=list(tt), optimisation_s=run.timings["optimisation_s"])), flush=True)
PY
cat $O/fit_cfg4share.txt | cut -c1-600
echo "== final default bench line"
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; cut -c1-250 $O/bench_default.json
ls $O
