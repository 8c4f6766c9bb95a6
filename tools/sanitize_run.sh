#!/bin/bash
TAG=${1:-r02}
mkdir -p gpurun_out
# compute-sanitizer passes: memcheck over the smoke (8 bond updates: DMMA kernels, fat kernel, SVD
# chain) and over the tcgen05 kernels (tools/oz_test: TMA / TMEM / mbarrier paths); racecheck over the smoke last
# (the QR / Jacobi kernels synchronise through shared-memory flags on purpose: hazards reported there are listed, not errors)
SAN=/usr/local/cuda/bin/compute-sanitizer
{
  echo "# compute-sanitizer, ${TAG}: $(${SAN} --version | head -2 | tr '\n' ' ')"
  echo "## memcheck: python __graft_entry__.py smoke"
  timeout 300 ${SAN} --tool memcheck --print-limit 20 python __graft_entry__.py smoke 2>&1 | grep -v '^$' | tail -12
  echo "## memcheck: tools/oz_test 2 (4096 x 128, S=4) and tools/oz_test 6 (20000 rows, K=100, S=2, div=10: ragged everything)"
  timeout 200 ${SAN} --tool memcheck --print-limit 20 tools/oz_test 2 2>&1 | grep -v '^$' | tail -8
  timeout 200 ${SAN} --tool memcheck --print-limit 20 tools/oz_test 6 2>&1 | grep -v '^$' | tail -8
  echo "## racecheck: python __graft_entry__.py smoke"
  timeout 240 ${SAN} --tool racecheck --print-limit 10 python __graft_entry__.py smoke 2>&1 | grep -v '^$' | tail -14
} > gpurun_out/sanitizer_${TAG}.txt 2>&1
