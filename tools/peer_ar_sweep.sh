#!/bin/bash
# Developer sweep of the peer all-reduce kernel (tuning build): blocks x {full, no handshakes, handshakes only}
# usage: tools/peer_ar_sweep.sh <nproc> <out file>
N=$1; OUT=$2; : > $OUT
for v in 3296 3148 3074 3037 4296 4148 4074 5296 5148; do
  echo "variant $v" >> $OUT
  B200SPLAT_TUNING_VARIANT=$v AR_ONLY=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tools/peer_bench.py 2>/dev/null | tail -1 >> $OUT
done
