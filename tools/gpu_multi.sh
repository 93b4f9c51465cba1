#!/bin/bash
# Multi-GPU check + bench (run under gpurun --gpus N): tools/gpu_multi.sh N TAG
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
{
for tr in p2p nccl; do
  for direct in 1 0; do
    [ "$tr" = nccl ] && [ "$direct" = 0 ] && continue
    echo "== shard check transport=$tr direct_out=$direct"
    CHECK_TRANSPORT=$tr CHECK_DIRECT=$direct CHECK_B=37 CHECK_LF=50 CHECK_MB=8 timeout 300 $TR tools/shard_nccl_check.py 2>&1 | grep -v "^\[W\|NCCL INFO\|^$" | tail -4
    CHECK_TRANSPORT=$tr CHECK_DIRECT=$direct CHECK_B=256 CHECK_LF=200 CHECK_MB=32 timeout 300 $TR tools/shard_nccl_check.py 2>&1 | grep -v "^\[W\|NCCL INFO\|^$" | tail -4
  done
done
echo "== bench --gpus $N"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "rc=$?"
tail -c 3000 gpurun_out/bench_n${N}_$TAG.json
tail -5 gpurun_out/bench_n${N}_$TAG.err
} 2>&1 | tee gpurun_out/multi_n${N}_$TAG.log
