#!/bin/bash
# 2-GPU exploration of the data-parallel overhead (one JSON line per configuration -> gpurun_out/ddp_explore.txt)
N=${1:-2}
run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-profile-gemm $EXTRA 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*' | head -2 | tr '\n' ' '; echo; }
EXTRA="" run A=1
EXTRA="" run NFB_DDP_DEBUG_SKIP_ALLREDUCE=1
EXTRA="" run NCCL_MAX_CTAS=4
EXTRA="" run NCCL_MAX_CTAS=8
EXTRA="" run NCCL_MIN_CTAS=32
EXTRA="--bucket-mb 50" run A=1
EXTRA="--bucket-mb 100" run A=1
EXTRA="--bucket-mb 12" run A=1
EXTRA="--no-wgrad-overlap" run A=1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-profile-gemm > gpurun_out/nccl_info.log 2>&1
grep -i "nvls\|algo\|channels\|nccl version" gpurun_out/nccl_info.log | sort | uniq -c | sort -rn | head -20
