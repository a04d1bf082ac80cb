#!/bin/bash
N=${1:-4}
run() { echo "== $* $EXTRA"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-profile-gemm $EXTRA 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*\|Error.*\|error.*' | head -3 | tr '\n' ' '; echo; }
EXTRA="" run A=1
EXTRA="" run NCCL_ALGO=NVLS
EXTRA="" run NCCL_ALGO=NVLSTree
EXTRA="" run NCCL_ALGO=Tree
EXTRA="" run NCCL_MAX_CTAS=16
EXTRA="" run NCCL_MAX_CTAS=24
EXTRA="--bucket-mb 50" run A=1
EXTRA="--bucket-mb 100" run A=1
EXTRA="--bucket-mb 100" run NCCL_ALGO=NVLS
EXTRA="--bucket-mb 50" run NCCL_MAX_CTAS=16
