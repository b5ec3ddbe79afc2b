#!/bin/bash
# tools/run_8gpu.sh [N] -- the three closed-loop BASELINE configurations on N GPUs of one node (default 8), one rank per GPU
N=${1:-8}
mkdir -p gpurun_out
for wl in tracking obstacles timeopt; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --workload $wl > gpurun_out/r2_bench_${wl}_${N}gpu.json 2> gpurun_out/r2_bench_${wl}_${N}gpu.err
  echo "$wl rc=$?"; cut -c1-400 gpurun_out/r2_bench_${wl}_${N}gpu.json; tail -3 gpurun_out/r2_bench_${wl}_${N}gpu.err
done
