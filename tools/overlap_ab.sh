#!/bin/bash
# 2-GPU session: DDP / FineTuner equivalence check, then the fine-tune step at both shapes; bench.py itself alternates the overlapped and
# the serial gradient exchange in-process (config.allreduce_overlap_ab_ms_per_step).
mkdir -p gpurun_out
TAG=${1:-r02z}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/ddp_check.py > gpurun_out/ddp_check_${TAG}.log 2>&1; echo "ddp_check rc=$?"; tail -2 gpurun_out/ddp_check_${TAG}.log
for wl in cfg2 cfg3; do
  timeout 900 $TR bench.py --gpus 2 --mode train --workload $wl --steps 5 --warmup 3 > gpurun_out/train_${wl}_n2_${TAG}.json 2> gpurun_out/train_${wl}_n2_${TAG}.err
  echo "train $wl rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/train_${wl}_n2_${TAG}.json')); c=d['config']
print(d['value'], d['ms_per_step'], 'payload', c['allreduce_payload_bytes'], 'alone', c['allreduce_alone_ms'], 'exposed', c['allreduce_exposed_ms'], 'A/B', c['allreduce_overlap_ab_ms_per_step'], c['gemm_route'])"
done
timeout 900 $TR bench.py --gpus 2 --no-secondary --no-cpu-baseline > gpurun_out/bench_n2_${TAG}.json 2> gpurun_out/bench_n2_${TAG}.err; echo "bench n2 rc=$?"
python tools/bench_brief.py gpurun_out/bench_n2_${TAG}.json | head -2 | cut -c1-250
