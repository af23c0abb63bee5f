#!/bin/bash
# r02f (2 GPUs): the whole GPU test suite (incl. handout images, multi-GPU library tests), strong scaling with NVLS off
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/handout_parity.jsonl $O/parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r02f_tests.log 2>&1; echo "tests rc=$?" >> $O/r02f_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02f_bench_2gpu.json 2> $O/r02f_bench_2gpu.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02f_smoke.log 2>&1
# the CLI end to end: XML -> image on 1 and 2 GPUs
cd /tmp && timeout 120 /root/repo/lajolla_public_b200/lajolla --spp 64 -o /root/repo/gpurun_out/r02f_cbox.pfm /root/repo/oracle/_ref/scenes/cbox/cbox.xml > /root/repo/gpurun_out/r02f_cli.log 2>&1
timeout 120 /root/repo/lajolla_public_b200/lajolla --spp 64 --gpus 2 -o /root/repo/gpurun_out/r02f_cbox_2gpu.exr /root/repo/oracle/_ref/scenes/cbox/cbox.xml >> /root/repo/gpurun_out/r02f_cli.log 2>&1
