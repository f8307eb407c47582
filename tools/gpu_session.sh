#!/bin/bash
# One GPU-box session: run the steps named on the command line, each with its own log under gpurun_out/.
# usage: tools/gpu_session.sh <tag> step1 step2 ...      (steps: tests parity bench bench_eager prof kbench ncu_launches)
tag=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    tests)   timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_configs.py > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" ;;
    alltests) timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${tag}_alltests.log 2>&1; echo "alltests rc=$?" ;;
    parity)  timeout 2400 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -s > gpurun_out/${tag}_parity.log 2>&1; echo "parity rc=$?" ;;
    bench)   timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/${tag}_prof.json > gpurun_out/${tag}_bench.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/${tag}_bench.log ;;
    smoke)   timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    *)       timeout 1200 bash -c "$step" > gpurun_out/${tag}_$(echo "$step" | tr -c 'a-zA-Z0-9' '_' | cut -c1-40).log 2>&1; echo "[$step] rc=$?" ;;
  esac
done
