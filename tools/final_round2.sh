# End-of-round evidence on one GPU: full GPU test suite (twice), smoke, the bench line, the reference arm, the other
# BASELINE configs and two micro-benchmarks.  Outputs under gpurun_out/final_*.
set -x
for i in 1 2; do
  timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py > gpurun_out/final_tests_$i.log 2>&1; echo "tests run $i rc=$?"
  tail -2 gpurun_out/final_tests_$i.log | cut -c1-200
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/final_prof.json > gpurun_out/final_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.log 2>&1; echo "reference rc=$?"
timeout 900 python tools/bench_configs.py --steps 10 --warmup 4 --fused-only > gpurun_out/final_bench_configs.jsonl 2> gpurun_out/final_bench_configs.err; echo "configs rc=$?"
python tools/kbench.py sn 128 1152 > gpurun_out/final_kbench_sn.txt 2>&1
python tools/kbench.py sn 16 144 >> gpurun_out/final_kbench_sn.txt 2>&1
python tools/kbench.py sn 256 2304 >> gpurun_out/final_kbench_sn.txt 2>&1
