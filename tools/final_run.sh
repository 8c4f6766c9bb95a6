#!/bin/bash
# Round-end measurement pass on the GPU box (one GPU): tests, bench line, BASELINE configs, full-sweep
# diagnostics, ncu evidence.  Outputs land in gpurun_out/ (scratch) and profiles/ (tracked).
TAG=${1:-r02}
export TNML_PROFILE_TAG=${TAG}
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/final_pytest.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
rm -f gpurun_out/configs_${TAG}.txt
timeout 400 python tools/run_configs.py 1 300 2 3 > gpurun_out/configs.log 2>&1
rm -f gpurun_out/sweep_diag.txt
timeout 300 python tools/sweep_diag.py 60000 120 4 > gpurun_out/diag_final.log 2>&1
cp gpurun_out/sweep_diag.txt gpurun_out/sweeps_${TAG}.txt
TNML_QR_DEBUG=1 timeout 120 python tools/svd_bench.py > gpurun_out/svd_bench_${TAG}.log 2>&1
timeout 120 tools/umma_bench > gpurun_out/umma_bench_${TAG}.txt 2>&1
timeout 200 tools/oz_test > gpurun_out/oz_test_${TAG}.txt 2>&1
bash tools/profile_run.sh ${TAG} > gpurun_out/profile_run.log 2>&1
bash tools/sanitize_run.sh ${TAG}
cat gpurun_out/final_pytest.log
head -c 600 gpurun_out/bench_final.json; echo
tail -2 gpurun_out/profile_run.log
