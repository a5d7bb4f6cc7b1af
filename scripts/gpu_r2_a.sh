#!/bin/bash
# round 2, GPU call A: the whole GPU suite (new BASELINE-size parity tests), default bench, batch-size sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
python bench.py > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err
for cfg in "592 1184 16" "1184 2368 32" "2368 4736 64" "4736 9472 96"; do
  set -- $cfg
  NCB_BENCH_PREWARM=2 NCB_BATCH=$1 NCB_BATCH_MEM_GB=$3 python bench.py --traj-per-step $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_b$1.json 2> gpurun_out/a_bench_b$1.err
done
tail -3 gpurun_out/a_pytest.log
cat gpurun_out/a_bench_*.json | cut -c1-400
