#!/bin/bash
# 2-GPU session: DDP / FineTuner equivalence check, then the fine-tune step with the gradient all-reduce overlapped (default) and serial.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/ddp_check.py > gpurun_out/ddp_check_r02x.log 2>&1; echo "ddp_check rc=$?"; tail -3 gpurun_out/ddp_check_r02x.log
for wl in cfg2 cfg3; do
  for ov in 1 0; do
    TCAVP_FT_OVERLAP=$ov timeout 900 $TR bench.py --gpus 2 --mode train --workload $wl --steps 5 --warmup 3 > gpurun_out/train_${wl}_n2_ov${ov}.json 2> gpurun_out/train_${wl}_n2_ov${ov}.err
    echo "train $wl overlap=$ov rc=$?"
    python -c "
import json; d=json.load(open('gpurun_out/train_${wl}_n2_ov${ov}.json')); c=d['config']
print(d['value'], d['ms_per_step'], 'payload', c['allreduce_payload_bytes'], 'alone', c['allreduce_alone_ms'], 'exposed', c['allreduce_exposed_ms'])"
  done
done
