#!/bin/bash
# Round-2 multi-GPU batch (run with gpurun --gpus G).  usage: bash scripts/gpu_r2_multi.sh G
#   G = 2: the 2-rank NCCL tests (sharded == single, lock-step fit, single-process launcher), N=2 scaling points,
#          reference arm under torchrun
#   G = 4: N=4 scaling points
#   G = 8: N=8 scaling points, BASELINE configs[3] (N=500k, L=500, 220 MB all-reduce), Hamming config 3 on 8 GPUs,
#          single-process run_plmc through the launcher at config-4 scale
cd "$(dirname "$0")/.."
G=${1:-8}
O=gpurun_out/r2m; mkdir -p $O
PORT=$((29700 + G * 10))
tr() { n=$1; shift; PORT=$((PORT+1)); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $n "$@"; }
show() { python -c "import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'value %.4g' % d['value'], d.get('rank_consistency'), d.get('communication'), d['roofline']['stage_ms'])" $1; }
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus_$G.txt
if [ $G -eq 2 ]; then
  echo "== multi-GPU tests"
  timeout 1500 python -m pytest tests/test_gpu_multi.py -q > $O/pytest_multi.txt 2>&1; echo "rc=$?"; tail -8 $O/pytest_multi.txt
  echo "== reference arm under torchrun (rank 0 only, all host threads)"
  tr 2 --impl reference --steps 2 --warmup 1 > $O/reference_torchrun_2.json 2>/dev/null; cut -c1-500 $O/reference_torchrun_2.json
  echo "== 1-GPU baseline of this box"
  timeout 600 python bench.py --steps 100 --no-subrecords > $O/weak_1.json 2> $O/weak_1.err; show $O/weak_1.json
fi
echo "== weak scaling point N=$G (50k sequences per GPU)"
tr $G --steps 100 --no-subrecords > $O/weak_$G.json 2> $O/weak_$G.err; show $O/weak_$G.json
echo "== strong scaling point N=$G (50k sequences in total, the size the metric is quoted on)"
tr $G --steps 100 --scaling strong --no-subrecords > $O/strong_$G.json 2> $O/strong_$G.err; show $O/strong_$G.json
grep -h "NCCL INFO.*\(nranks\|NVLS\|Connected\)" $O/weak_$G.err | head -6
echo "== bf16 tiles, weak, N=$G"
tr $G --steps 100 --precision bf16 --no-subrecords > $O/weak_bf16_$G.json 2> /dev/null; show $O/weak_bf16_$G.json
if [ $G -eq 8 ]; then
  echo "== BASELINE configs[3]: N=500,000 L=500 sharded over 8 GPUs, one all-reduce of 220 MB per evaluation"
  tr 8 --seqs 62500 --sites 500 --steps 10 --no-subrecords > $O/cfg4_8.json 2> $O/cfg4_8.err; echo "rc=$?"; show $O/cfg4_8.json
  echo "== all-reduce time vs message size (transfer vs waiting)"
  PORT=$((PORT+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT scripts/nccl_sizes.py > $O/nccl_sizes_8.json 2>/dev/null; tail -1 $O/nccl_sizes_8.json | cut -c1-1500
  echo "== Hamming config 3 on 8 GPUs"
  PORT=$((PORT+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --workload hamming --steps 5 > $O/hamming_8.json 2>/dev/null; cut -c1-300 $O/hamming_8.json
  echo "== single-process run_plmc over 8 GPUs (launcher), config-4 alignment written as A2M"
  timeout 1200 python - > $O/launcher_cfg4.txt 2>&1 <<PY
import os, sys, time, json
sys.path.insert(0, os.getcwd())
from evcouplings_b200 import synthetic, tools
N, L = 500000, 500
t0 = time.time(); codes = synthetic.synthetic_msa_codes(N, L, 4); a2m = "/tmp/cfg4.a2m"; synthetic.write_a2m(a2m, codes)
print("wrote A2M", N, L, "in %.1f s" % (time.time() - t0), flush=True)
t0 = time.time()
res, run = tools.run_plmc(a2m, "/tmp/cfg4_ECs.txt", "/tmp/cfg4.model", focus_seq="seq0", theta=0.8, iterations=20,
                          lambda_h=0.01, lambda_J=0.01 * 20 * (L - 1), cpu=8, return_run=True)
wall = time.time() - t0
fx = res.iteration_table["fx"].astype(float).values
tt = res.iteration_table["time"].astype(float).values
print("iteration-table time column (s):", list(tt[:4]), "...", list(tt[-2:]), flush=True)
print(json.dumps(dict(wall_s=wall, timings=run.timings, iterations=len(fx), fx=list(fx[:3]) + list(fx[-2:]),
                      status=res.optimization_status, n_eff=res.effective_samples,
                      model_bytes=os.path.getsize("/tmp/cfg4.model"))), flush=True)
PY
  tail -3 $O/launcher_cfg4.txt
fi
ls $O
