#!/bin/bash
# Round-2 multi-GPU batch (run with gpurun --gpus G): 2-rank NCCL tests, weak / strong scaling lines,
# BASELINE configs[3] (N=500k, L=500) sharded over all GPUs, single-process run_plmc through the launcher,
# Hamming config 3 on all GPUs.   usage: bash scripts/gpu_r2_multi.sh G
cd "$(dirname "$0")/.."
G=${1:-8}
O=gpurun_out/r2m; mkdir -p $O
PORT=29700
tr() { n=$1; shift; PORT=$((PORT+1)); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $n "$@"; }
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
echo "== multi-GPU tests"
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -x > $O/pytest_multi.txt 2>&1; echo "rc=$?"; tail -5 $O/pytest_multi.txt
echo "== weak scaling (50k sequences per GPU)"
for n in 1 2 4 8; do [ $n -le $G ] || continue
  if [ $n -eq 1 ]; then timeout 600 python bench.py --steps 100 --no-subrecords > $O/weak_$n.json 2> $O/weak_$n.err
  else tr $n --steps 100 --no-subrecords > $O/weak_$n.json 2> $O/weak_$n.err; fi
  python -c "import json; d=json.load(open('$O/weak_$n.json')); print('weak', $n, d['ms_per_step'], d['value'], d.get('rank_consistency'), d.get('accuracy'))"
done
echo "== strong scaling (N=50k total, the size the metric is quoted on)"
for n in 2 4 8; do [ $n -le $G ] || continue
  tr $n --steps 100 --scaling strong --no-subrecords > $O/strong_$n.json 2> $O/strong_$n.err
  python -c "import json; d=json.load(open('$O/strong_$n.json')); print('strong', $n, d['ms_per_step'], d['value'], d['roofline']['stage_ms'])"
done
echo "== bf16 tiles, weak, all GPUs"
tr $G --steps 100 --precision bf16 --no-subrecords > $O/weak_bf16_$G.json 2> /dev/null; python -c "import json; d=json.load(open('$O/weak_bf16_$G.json')); print('weak bf16', $G, d['ms_per_step'], d['value'])"
echo "== BASELINE configs[3]: N=62500 x $G sequences, L=500, one all-reduce of 220 MB per evaluation"
tr $G --seqs 62500 --sites 500 --steps 10 --no-subrecords > $O/cfg4_$G.json 2> $O/cfg4_$G.err; echo "rc=$?"
python -c "import json; d=json.load(open('$O/cfg4_$G.json')); print('cfg4', $G, d['ms_per_step'], d['value'], d['config']['parallelism'], d.get('rank_consistency'), d.get('accuracy'))"
echo "== reference arm under torchrun (rank 0 only, all host threads)"
tr $G --impl reference --steps 2 --warmup 1 > $O/reference_torchrun_$G.json 2>/dev/null; cut -c1-400 $O/reference_torchrun_$G.json
echo "== Hamming config 3 on $G GPUs"
PORT=$((PORT+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $G --workload hamming --steps 5 > $O/hamming_$G.json 2>/dev/null; cut -c1-300 $O/hamming_$G.json
echo "== single-process run_plmc over all GPUs (launcher), config-4-like alignment written as A2M"
timeout 1500 python - > $O/launcher_cfg4.txt 2>&1 <<PY
import os, sys, time, json
sys.path.insert(0, os.getcwd())
from evcouplings_b200 import synthetic, tools
N, L = 62500 * $G, 500
t0 = time.time(); codes = synthetic.synthetic_msa_codes(N, L, 4); a2m = "/tmp/cfg4.a2m"; synthetic.write_a2m(a2m, codes)
print("wrote A2M", N, L, "in %.1f s" % (time.time() - t0), flush=True)
out = {}
for ng, iters in (($G, 20), (1, 5)):
    t0 = time.time()
    res, run = tools.run_plmc(a2m, "/tmp/cfg4_%d_ECs.txt" % ng, "/tmp/cfg4_%d.model" % ng, focus_seq="seq0", theta=0.8,
                              iterations=iters, lambda_h=0.01, lambda_J=0.01 * 20 * (L - 1), num_gpus=ng, return_run=True)
    wall = time.time() - t0
    fx = res.iteration_table["fx"].astype(float).values
    out[ng] = dict(wall_s=wall, timings=run.timings, iterations=len(fx), fx_first5=list(fx[:5]), status=res.optimization_status,
                   n_eff=res.effective_samples)
    print(json.dumps({ng: out[ng]}), flush=True)
a, b = out[$G]["fx_first5"], out[1]["fx_first5"]
print("fx agreement multi vs single GPU (first 5 iterations): max rel diff %.2e" % max(abs(x - y) / abs(y) for x, y in zip(a, b)))
PY
tail -5 $O/launcher_cfg4.txt
ls -la $O
