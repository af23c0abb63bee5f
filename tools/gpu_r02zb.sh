#!/bin/bash
# r02zb (1 GPU): winner primitive fetched once at write-back; queue-form extension for volpath scenes with a hierarchy; A/B against the previous build
mkdir -p gpurun_out; O=gpurun_out/r02zb_ab.txt; : > $O
run() {
  python bench.py --workload $1 --steps 2 --warmup 2 --spp 256 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$1 $2', round(j['value'], 1), {k: round(v, 1) for k, v in st.items()})" >> $O 2>&1
}
for W in sponza vol_cbox_teapot disney_bsdf; do
  LJ_LIB=$PWD/lajolla_public_b200/build/libljb200_prev.so run $W "previous build"
  run $W "this build"
done
timeout 600 python -m pytest tests -m gpu -q -k "ray_parity or wavefront or walk or teapot" > gpurun_out/r02zb_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02zb_tests.log
