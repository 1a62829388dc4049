#!/bin/bash
# One GPU-box session: parity tests, default bench (+ reference arm), cfg3 bench, launch list + full ncu capture.
TAG=${1:-r01d}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_${TAG}.json
python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_${TAG}.json 2> gpurun_out/bench_cfg3_${TAG}.err; echo "cfg3 rc=$?"; tail -3 gpurun_out/bench_cfg3_${TAG}.err
python tools/bench_brief.py gpurun_out/bench_${TAG}.json gpurun_out/bench_cfg3_${TAG}.json 2>&1 | cut -c1-330
SKIP=40 COUNT=3 tools/profile.sh ${TAG} "gemm_tc_kernel<256"
