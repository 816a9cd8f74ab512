#!/bin/bash
# Round-2 GPU batch D (1 GPU): lean single-lane issue (elect.sync) for the MMA / TMA warps + SINGLE as a template
# parameter; then the full regression (all GPU tests, bench lines of both precision modes and the big shapes, ncu).
cd "$(dirname "$0")/.."
O=gpurun_out/r2d; mkdir -p $O
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; rc=$?; tail -2 $O/smoke.txt
if [ $rc -ne 0 ]; then echo "SMOKE FAILED rc=$rc"; tail -30 $O/smoke.txt; exit 1; fi
for sp in 0 1 2; do
  for prec in fp32 bf16; do
    EVC_SPLIT_PRODUCER=$sp timeout 300 python bench.py --no-subrecords --steps 60 --precision $prec > $O/bench_sp${sp}_$prec.json 2>/dev/null
    python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],4), {k[:12]: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, d['accuracy']['grad_rel_l2_err'])" $O/bench_sp${sp}_$prec.json
  done
done
echo "== full GPU test suite"
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_all.txt 2>&1; echo "rc=$?"; tail -6 $O/pytest_all.txt
echo "== bench lines"
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; cut -c1-300 $O/bench_default.json
timeout 600 python bench.py --precision bf16 --no-subrecords > $O/bench_bf16.json 2>/dev/null; cut -c1-300 $O/bench_bf16.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-subrecords > $O/bench_20steps.json 2>/dev/null; cut -c1-300 $O/bench_20steps.json
timeout 900 python bench.py --seqs 62500 --sites 500 --steps 20 --no-subrecords > $O/bench_cfg4share.json 2>/dev/null; cut -c1-200 $O/bench_cfg4share.json
timeout 900 python bench.py --seqs 100000 --sites 800 --steps 10 --no-subrecords > $O/bench_cfg5_fp32.json 2>/dev/null; cut -c1-200 $O/bench_cfg5_fp32.json
timeout 900 python bench.py --seqs 100000 --sites 800 --steps 10 --no-subrecords --precision bf16 > $O/bench_cfg5_bf16.json 2>/dev/null; cut -c1-200 $O/bench_cfg5_bf16.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null; cut -c1-200 $O/bench_reference.json
echo "== ncu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_fp32.csv python bench.py --steps 3 --warmup 3 --no-subrecords > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_persistent|plm_softmax" -s 6 -c 3 -o $O/prof_fp32 python bench.py --steps 2 --warmup 3 --no-subrecords > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_persistent|plm_softmax" -s 6 -c 3 -o $O/prof_bf16 python bench.py --steps 2 --warmup 3 --no-subrecords --precision bf16 > /dev/null 2>&1
ls $O
