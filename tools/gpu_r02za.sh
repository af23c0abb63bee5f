#!/bin/bash
# r02za (1 GPU): Lambertian-only shade kernel (scenes whose every material is Lambertian), 4 / 5 / 6 resident CTAs per SM
mkdir -p gpurun_out; O=gpurun_out/r02za_lambert.txt; : > $O
run() {
  python bench.py --workload $1 --steps 2 --warmup 2 --spp 256 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$1 $2', round(j['value'], 1), {k: round(v, 1) for k, v in st.items()})" >> $O 2>&1
}
for W in sponza cbox; do
  LJ_LAMBERT_KERNEL=0 run $W "general kernel"
  run $W "lambert, 4 CTAs/SM"
  LJ_LIB=$PWD/lajolla_public_b200/build/libljb200_l5.so run $W "lambert, 5 CTAs/SM"
  LJ_LIB=$PWD/lajolla_public_b200/build/libljb200_l6.so run $W "lambert, 6 CTAs/SM"
done
timeout 600 python -m pytest tests -m gpu -q -k "image_parity and (cbox or sponza) or golden_tiles or additive or full_size" > gpurun_out/r02za_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02za_tests.log
