#!/bin/bash
# bench.py at N GPUs with NCCL confined to fewer CTAs (the exchange overlaps the next step's trace and takes its SMs)
N=${1:-2}
for ctas in default 2 4 8; do
  if [ "$ctas" = default ]; then unset NCCL_MAX_CTAS NCCL_MIN_CTAS; else export NCCL_MAX_CTAS=$ctas NCCL_MIN_CTAS=1; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py \
      --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('NCCL_MAX_CTAS=$ctas', 'value %.4e' % d['value'], 'ms_per_step %.3f' % d['ms_per_step'], 'kernel_ms', ['%.3f' % x for x in d['roofline']['kernel_ms_per_rank']])
"
done
