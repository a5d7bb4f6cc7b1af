#!/bin/bash
# round 2, 2-GPU call: the N > 1 arms of bench.py (NCCL all-reduce, backend.execute_sharded, strong scaling, reference arm)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 6 --warmup 4 > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err
echo "exit $?"; tail -c 1500 gpurun_out/g2_bench.json; tail -5 gpurun_out/g2_bench.err
$TR bench.py --impl reference --gpus 2 --steps 2 --warmup 0 > gpurun_out/g2_ref.json 2> gpurun_out/g2_ref.err
echo "exit $?"; cut -c1-300 gpurun_out/g2_ref.json
