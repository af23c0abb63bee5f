#!/bin/bash
# r02z (1 GPU): render buffers recycled across scenes (close / create without cudaFree): e2e check + regression tests
mkdir -p gpurun_out; O=gpurun_out
python tools/diag_e2e2.py sponza 256 0 > $O/r02z_diag_e2e.txt 2>&1
python tools/diag_e2e2.py cbox 64 0 >> $O/r02z_diag_e2e.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "abi or image_parity or additive or aux or full_size or frontend or cli" > $O/r02z_tests.log 2>&1; echo "tests rc=$?" >> $O/r02z_tests.log
for W in sponza cbox disney_bsdf; do
  timeout 500 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > $O/r02z_bench_$W.json 2> $O/r02z_bench_$W.err
done
