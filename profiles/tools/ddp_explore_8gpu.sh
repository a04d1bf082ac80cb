#!/bin/bash
N=8
run() { echo "== $* $EXTRA"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-profile-gemm $EXTRA 2>&1 | grep -o '"ms_per_step": [0-9.]*\|B200Error: .*' | head -1 | cut -c1-160; }
EXTRA="--bucket-mb 25" run A=1
EXTRA="--bucket-mb 100" run A=1
EXTRA="--bucket-mb 50" run "NCCL_ALGO=allreduce:NVLS"
EXTRA="--bucket-mb 50" run "NCCL_ALGO=allreduce:Ring"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=TUNING python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --no-profile-gemm 2>&1 | grep "AllReduce:" | sed 's/.*AllReduce: //' | sort | uniq -c | sort -rn | head -5
