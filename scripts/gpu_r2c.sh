#!/bin/bash
# Round-2 GPU batch C1 (1 GPU): number of TMA producer threads per CTA (EVC_SPLIT_PRODUCER = 0 / 1 / 2), both precisions
cd "$(dirname "$0")/.."
O=gpurun_out/r2c; mkdir -p $O
for sp in 2 1 0; do
  echo "== EVC_SPLIT_PRODUCER=$sp"
  EVC_SPLIT_PRODUCER=$sp timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_sp$sp.txt 2>&1 || { echo "smoke failed"; tail -5 $O/smoke_sp$sp.txt; continue; }
  for prec in fp32 bf16; do
    EVC_SPLIT_PRODUCER=$sp timeout 300 python bench.py --no-subrecords --steps 60 --precision $prec > $O/bench_sp${sp}_$prec.json 2>/dev/null
    python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],4), {k[:12]: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, d['accuracy']['grad_rel_l2_err'])" $O/bench_sp${sp}_$prec.json
  done
done
EVC_SPLIT_PRODUCER=2 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_core or large_L or full_size_properties or precision_bf16" > $O/pytest_sp2.txt 2>&1; echo "rc=$?"; tail -3 $O/pytest_sp2.txt
EVC_SPLIT_PRODUCER=2 timeout 600 python bench.py --no-subrecords --steps 10 --seqs 100000 --sites 800 --precision bf16 > $O/bench_sp2_cfg5_bf16.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['roofline']['stage_ms'], d['clocks'])" $O/bench_sp2_cfg5_bf16.json
timeout 600 ncu --set full --clock-control none -k regex:"tc_gemm_persistent" -s 6 -c 2 -o $O/prof_sp2_bf16 env EVC_SPLIT_PRODUCER=2 python bench.py --steps 2 --warmup 3 --no-subrecords --precision bf16 > /dev/null 2>&1
ls $O
