#!/bin/bash
# r02zf (1 GPU): final validation of the session: full GPU suite, smoke, headline bench with the CPU baseline, the other configs without it
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/handout_parity.jsonl $O/parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r02zf_tests.log 2>&1; echo "tests rc=$?" >> $O/r02zf_tests.log
cp $O/parity.jsonl $O/r02zf_parity.jsonl; cp $O/handout_parity.jsonl $O/r02zf_handout_parity.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02zf_smoke.log 2>&1
timeout 500 python bench.py > $O/r02zf_bench_sponza.json 2> $O/r02zf_bench_sponza.err
for W in cbox veach_mi disney_bsdf volpath_test6 vol_cbox_teapot hetvol hetvol_colored; do
  timeout 300 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > $O/r02zf_bench_$W.json 2> $O/r02zf_bench_$W.err
done
