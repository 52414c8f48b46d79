#!/bin/bash
# 8-GPU box: why do 4 torchrun ranks get half the per-GPU host->device rate?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 tools/n4_probe.py "${@:3}" 2>/dev/null | grep '^{' | python -c "
import sys, json
r = sorted((json.loads(l) for l in sys.stdin), key=lambda d: d['rank'])
print('   per rank GB/s:', [d['gbs'] for d in r], 'sum', round(sum(d['gbs'][-1] for d in r), 1), 'omp', r[0]['omp'] if r else None)"; }
{
echo "== tool, 4 processes, GPUs 0-3"; python tools/h2d_ceiling.py --procs --rows 4096 --width 110250 --pitch 2646016 --reps 24 0,1,2,3 | cut -c1-400
run 4 29611 --backend nccl --torch-cuda 1
run 4 29612 --backend gloo --torch-cuda 1
run 4 29613 --backend gloo --torch-cuda 0
run 4 29614 --backend nccl --torch-cuda 1 --big-device-buffer 1
run 2 29615 --backend nccl --torch-cuda 1 --big-device-buffer 1
run 8 29616 --backend nccl --torch-cuda 1
echo "== 4 ranks on GPUs 4-7 (CUDA_VISIBLE_DEVICES)"; CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29617 tools/n4_probe.py --backend nccl 2>/dev/null | grep '^{' | cut -c1-200
echo "== 4 ranks, NCCL_P2P_DISABLE / SHM_DISABLE"; NCCL_P2P_DISABLE=1 NCCL_SHM_DISABLE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29618 tools/n4_probe.py --backend nccl 2>/dev/null | grep '^{' | cut -c1-200
} > gpurun_out/r2n4b.txt 2>&1
cat gpurun_out/r2n4b.txt
