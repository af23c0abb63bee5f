#!/bin/bash
# r02y (1 GPU): full suite + all eight workloads with the final build of this session; ncu launch list and metrics of sponza
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/handout_parity.jsonl $O/parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r02y_tests.log 2>&1; echo "tests rc=$?" >> $O/r02y_tests.log
for W in sponza cbox veach_mi disney_bsdf volpath_test6 vol_cbox_teapot hetvol hetvol_colored; do
  timeout 500 python bench.py --workload $W --steps 3 --warmup 3 > $O/r02y_bench_$W.json 2> $O/r02y_bench_$W.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02y_launches_sponza.csv python bench.py --steps 1 --warmup 1 --spp 64 --no-cpu-baseline > $O/r02y_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_trace_q|k_shade|k_regen' -s 24 -c 4 -o /tmp/r02y_sponza python bench.py --steps 1 --warmup 0 --spp 128 --no-cpu-baseline > $O/r02y_ncu.log 2>&1
python tools/ncu_metrics.py /tmp/r02y_sponza.ncu-rep > $O/r02y_sponza_metrics.txt 2>&1
