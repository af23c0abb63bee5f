#!/bin/bash
# r02zd (1 GPU): k_trace_q with a second node step in the same pass for rays that still want one; A/B
mkdir -p gpurun_out; O=gpurun_out/r02zd_ab.txt; : > $O
run() {
  python bench.py --workload $1 --steps 2 --warmup 2 --spp 256 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$1 $2', round(j['value'], 1), {k: round(v, 1) for k, v in st.items()})" >> $O 2>&1
}
for W in sponza disney_bsdf; do
  run $W "1 node step per pass"
  LJ_LIB=$PWD/lajolla_public_b200/build/libljb200_n2.so run $W "2 node steps per pass"
done
LJ_LIB=$PWD/lajolla_public_b200/build/libljb200_n2.so timeout 600 python -m pytest tests -m gpu -q -k "wavefront_kernels_ray_parity" > gpurun_out/r02zd_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02zd_tests.log
