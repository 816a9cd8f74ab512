#!/bin/bash
# Round-2 GPU batch F (2 GPUs): where does the time go in a multi-rank fit started by the launcher?
cd "$(dirname "$0")/.."
O=gpurun_out/r2f; mkdir -p $O
timeout 900 python - > $O/launcher_2gpu.txt 2>&1 <<PY
import os, sys, time, json
sys.path.insert(0, os.getcwd())
os.environ["EVC_TRACE"] = "1"
from evcouplings_b200 import synthetic, tools
N, L = 125000, 500
codes = synthetic.synthetic_msa_codes(N, L, 4); a2m = "/tmp/cfg4q.a2m"; synthetic.write_a2m(a2m, codes)
for ng in (2, 1):
    t0 = time.time()
    res, run = tools.run_plmc(a2m, "/tmp/q%d_ECs.txt" % ng, "/tmp/q%d.model" % ng, focus_seq="seq0", theta=0.8, iterations=10,
                              lambda_h=0.01, lambda_J=0.01 * 20 * (L - 1), num_gpus=ng, return_run=True)
    tt = res.iteration_table["time"].astype(float).values
    print("GPUS", ng, json.dumps(dict(wall_s=time.time() - t0, timings=run.timings, time_column=list(tt))), flush=True)
PY
grep -v "evc-trace" $O/launcher_2gpu.txt | cut -c1-1200; grep "evc-trace" $O/launcher_2gpu.txt | head -40
