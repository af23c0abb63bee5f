#!/bin/bash
# r02zc (1 GPU): k_trace_q tests up to 2 / 3 primitives of a ray's group per pass; A/B against one per pass
mkdir -p gpurun_out; O=gpurun_out/r02zc_ab.txt; : > $O
run() {
  python bench.py --workload $1 --steps 2 --warmup 2 --spp 256 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$1 $2', round(j['value'], 1), {k: round(v, 1) for k, v in st.items()}, 'prim simd', round(j['simd_efficiency']['prim_step'], 3))" >> $O 2>&1
}
for W in sponza disney_bsdf; do
  LJ_LIB=$PWD/lajolla_public_b200/build/libljb200_prev.so run $W "1 per pass"
  run $W "2 per pass"
  LJ_LIB=$PWD/lajolla_public_b200/build/libljb200_p3.so run $W "3 per pass"
done
timeout 600 python -m pytest tests -m gpu -q -k "ray_parity or wavefront" > gpurun_out/r02zc_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02zc_tests.log
