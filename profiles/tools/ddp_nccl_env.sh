#!/bin/bash
# NCCL knobs at N GPUs with the bench defaults (bf16 buckets from 4 GPUs on)
N=${1:-8}
run() {
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    --no-profile --no-other-configs --no-ddp-parity 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), round(d['value']))" || echo "$* FAILED"
}
run NCCL_DEBUG=WARN
run NCCL_MAX_CTAS=24
run NCCL_MAX_CTAS=16
run NCCL_MAX_CTAS=8
run NCCL_ALGO=NVLS
run NCCL_PROTO=Simple
run NCCL_NVLS_ENABLE=1 NCCL_ALGO=NVLSTree
