#!/bin/bash
# One GPU-box session: parity tests, default bench (+ reference arm), cfg3 / cfg5 / train benches, launch list + full ncu captures.
TAG=${1:-r02a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_${TAG}.json
python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_${TAG}.json 2> gpurun_out/bench_cfg3_${TAG}.err; echo "cfg3 rc=$?"; tail -3 gpurun_out/bench_cfg3_${TAG}.err
python bench.py --workload cfg5 --steps 50 --warmup 5 --cpu-sample 256 > gpurun_out/bench_cfg5_${TAG}.json 2> gpurun_out/bench_cfg5_${TAG}.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/bench_cfg5_${TAG}.err
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/train_cfg2_${TAG}.json 2> gpurun_out/train_cfg2_${TAG}.err; echo "train rc=$?"
python bench.py --mode train --workload cfg3 --steps 3 --warmup 3 > gpurun_out/train_cfg3_${TAG}.json 2> gpurun_out/train_cfg3_${TAG}.err; echo "train cfg3 rc=$?"
python tools/bench_brief.py gpurun_out/bench_${TAG}.json gpurun_out/bench_cfg3_${TAG}.json gpurun_out/bench_cfg5_${TAG}.json 2>&1 | cut -c1-330 | grep -v "^    " 
python -c "
import json
for f in ('gpurun_out/train_cfg2_${TAG}.json','gpurun_out/train_cfg3_${TAG}.json'):
    d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['config']['peak_mem_gib'])
"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_${TAG}.log
[ "${PROFILE:-1}" = "1" ] || exit 0
TRAFFIC_KERNEL=gemm_tc_pair_kernel LIST=${LIST:-1} COUNT=${COUNT:-4} tools/profile.sh ${TAG} "gemm_tc_pair_kernel@316" "attn_flash_kernel<\(int\)64@40" "row_rstd_kernel@2"
