#!/bin/bash
# Round-2 GPU batch B (1 GPU): CTA-pair (cta_group::2) GEMM tiles behind EVC_TC_PAIR=1 -- parity first, then speed;
# then the regression items of batch A (precision-schedule test, default bench line with sub-records).
cd "$(dirname "$0")/.."
O=gpurun_out/r2b; mkdir -p $O
echo "== pair kernel: smoke-size parity (bounded waits trap instead of hanging)"
EVC_TC_PAIR=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/pair_smoke.txt 2>&1; rc=$?; tail -3 $O/pair_smoke.txt
if [ $rc -eq 0 ]; then
  echo "== pair kernel: parity tests"
  EVC_TC_PAIR=1 timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_core or large_L or full_size_properties or precision_bf16 or zero_and_large or all_gap" > $O/pair_pytest.txt 2>&1; echo "rc=$?"; tail -6 $O/pair_pytest.txt
  echo "== pair kernel: speed"
  for prec in fp32 bf16; do
    EVC_TC_PAIR=1 timeout 300 python bench.py --no-subrecords --steps 60 --precision $prec > $O/pair_bench_$prec.json 2>/dev/null
    python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['roofline']['stage_ms'], d.get('accuracy'))" $O/pair_bench_$prec.json
  done
  EVC_TC_PAIR=1 timeout 600 python bench.py --no-subrecords --steps 10 --seqs 100000 --sites 800 --precision bf16 > $O/pair_bench_cfg5_bf16.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['roofline']['stage_ms'], d['clocks'])" $O/pair_bench_cfg5_bf16.json
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_pair" -s 6 -c 2 -o $O/prof_pair_bf16 env EVC_TC_PAIR=1 python bench.py --steps 2 --warmup 3 --no-subrecords --precision bf16 > /dev/null 2>&1
else
  echo "PAIR SMOKE FAILED rc=$rc"; tail -20 $O/pair_smoke.txt
fi
echo "== regression: precision schedule test, default line with sub-records"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "precision_schedule" -s > $O/pytest_sched.txt 2>&1; echo "rc=$?"; tail -4 $O/pytest_sched.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; python -c "import json,sys; d=json.load(open(sys.argv[1])); print(d['ms_per_step'], d['roofline']['traffic'], d['fit'], d['clocks'])" $O/bench_default.json
timeout 600 python bench.py --workload hamming --hamming-pabp --steps 5 > $O/bench_hamming_pabp.json 2>/dev/null; cut -c1-400 $O/bench_hamming_pabp.json
ls $O
