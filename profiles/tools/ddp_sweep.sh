#!/bin/bash
# data-parallel experiments at N GPUs: bench.py under torchrun with different bucket dtypes / sizes / embedding-gradient exchange
N=${1:-2}
shift
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    --no-profile --no-other-configs --no-ddp-parity "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), round(d['value']))"
}
run --grad-dtype float32
run --grad-dtype bfloat16
run --grad-dtype bfloat16 --sparse-embedding-grads
run --grad-dtype float32 --sparse-embedding-grads
run --grad-dtype bfloat16 --bucket-mb 50
run --grad-dtype bfloat16 --bucket-mb 200
run --grad-dtype bfloat16 --bucket-mb 25
