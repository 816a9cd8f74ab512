#!/bin/bash
# Round-2 GPU batch A: smoke, GPU tests, bench lines (fp32 / bf16 / big shapes / reference arm), mgroup sweep, ncu.
cd "$(dirname "$0")/.."
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
echo "== sparse metadata probe"; timeout 120 ./build/sparse_probe > $O/sparse_probe.txt 2>&1; echo "rc=$?"; tail -12 $O/sparse_probe.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; rc=$?; tail -3 $O/smoke.txt
if [ $rc -ne 0 ]; then echo "SMOKE FAILED rc=$rc"; tail -30 $O/smoke.txt; exit 1; fi
echo "== new tests first"
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x -k "precision or device_fit or create_rejects or identities or iteration_capped or config1" -s > $O/pytest_new.txt 2>&1; echo "rc=$?"; tail -15 $O/pytest_new.txt
echo "== bench default"
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; cut -c1-600 $O/bench_default.json
echo "== bench bf16"
timeout 600 python bench.py --precision bf16 --no-subrecords > $O/bench_bf16.json 2> $O/bench_bf16.err; echo "rc=$?"; cut -c1-400 $O/bench_bf16.json
echo "== fused forward"
timeout 600 python bench.py --forward tcfused --no-subrecords --steps 100 > $O/bench_fused.json 2>/dev/null; cut -c1-300 $O/bench_fused.json
timeout 600 python bench.py --forward tcfused --precision bf16 --no-subrecords --steps 100 > $O/bench_fused_bf16.json 2>/dev/null; cut -c1-300 $O/bench_fused_bf16.json
echo "== mgroup sweep (forward tile order), fp32 and bf16"
for mb in 8 16 36 48 64; do
  EVC_MGROUP_MB=$mb timeout 300 python bench.py --no-subrecords --steps 60 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mb=$mb fp32', d['ms_per_step'], d['roofline']['stage_ms'])"
  EVC_MGROUP_MB=$mb timeout 300 python bench.py --no-subrecords --steps 60 --precision bf16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mb=$mb bf16', d['ms_per_step'], d['roofline']['stage_ms'])"
done > $O/mgroup_sweep.txt 2>&1; cat $O/mgroup_sweep.txt
echo "== big shapes"
timeout 900 python bench.py --seqs 62500 --sites 500 --steps 20 --no-subrecords > $O/bench_cfg4share.json 2>/dev/null; cut -c1-300 $O/bench_cfg4share.json
timeout 900 python bench.py --seqs 100000 --sites 800 --steps 10 --no-subrecords > $O/bench_cfg5_fp32.json 2>/dev/null; cut -c1-300 $O/bench_cfg5_fp32.json
timeout 900 python bench.py --seqs 100000 --sites 800 --steps 10 --no-subrecords --precision bf16 > $O/bench_cfg5_bf16.json 2>/dev/null; cut -c1-300 $O/bench_cfg5_bf16.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null; cut -c1-300 $O/bench_reference.json
echo "== ncu launch list + full capture"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_fp32.csv python bench.py --steps 3 --warmup 3 --no-subrecords > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_persistent|plm_softmax" -s 6 -c 3 -o $O/prof_fp32 python bench.py --steps 2 --warmup 3 --no-subrecords > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_persistent|plm_softmax" -s 6 -c 3 -o $O/prof_bf16 python bench.py --steps 2 --warmup 3 --no-subrecords --precision bf16 > /dev/null 2>&1
ls -la $O
echo "== full GPU test suite"
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_all.txt 2>&1; echo "rc=$?"; tail -25 $O/pytest_all.txt
