#!/bin/bash
# Round-2 GPU batch G (8 GPUs): the single-process run_plmc over 8 GPUs at config-4 scale, traced (EVC_TRACE)
cd "$(dirname "$0")/.."
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python - > $O/launcher_cfg4_8gpu.txt 2>&1 <<PY
import os, sys, time, json
sys.path.insert(0, os.getcwd())
os.environ["EVC_TRACE"] = "1"
from evcouplings_b200 import synthetic, tools
N, L = 500000, 500
t0 = time.time(); codes = synthetic.synthetic_msa_codes(N, L, 4); a2m = "/tmp/cfg4.a2m"; synthetic.write_a2m(a2m, codes)
print("wrote A2M", N, L, "in %.1f s" % (time.time() - t0), flush=True)
for it in (20, 100):
    t0 = time.time()
    res, run = tools.run_plmc(a2m, "/tmp/cfg4_ECs.txt", "/tmp/cfg4.model", focus_seq="seq0", theta=0.8, iterations=it,
                              lambda_h=0.01, lambda_J=0.01 * 20 * (L - 1), cpu=8, return_run=True)
    tt = res.iteration_table["time"].astype(float).values
    print("ITER", it, json.dumps(dict(wall_s=time.time() - t0, timings=run.timings, time_column=list(tt[:5]) + list(tt[-3:]),
          status=res.optimization_status, n_eff=res.effective_samples, model_bytes=os.path.getsize("/tmp/cfg4.model"))), flush=True)
PY
grep -v "evc-trace" $O/launcher_cfg4_8gpu.txt | cut -c1-1500; grep "rank 0\|pid" $O/launcher_cfg4_8gpu.txt | grep -v "rank [1-7]" | head -40
