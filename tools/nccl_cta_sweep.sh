#!/bin/bash
# hires leg at N GPUs under different caps on the CTAs NCCL may use (its kernels share the SMs with the density kernels
# they overlap): one JSON line per setting in gpurun_out/nccl_cta_sweep.jsonl
N=${1:-8}
O=gpurun_out/nccl_cta_sweep.jsonl
: > $O
for ctas in default 16 8 4; do
  if [ $ctas = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
  echo "NCCL_MAX_CTAS=$ctas" >> $O
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload hires --hires-steps 3 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'ms': d['value'], 'n': d['n_gpus']}))" >> $O
done
cat $O
