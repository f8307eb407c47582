python -m pytest tests/test_gpu_tc.py tests/test_gpu_paths.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
for cfg in "conv 256 128 128 16 16 3" "conv 256 64 64 32 32 3" "conv 256 32 32 64 64 3" "conv 256 16 16 128 128 3" "conv 256 64 64 16 32 3"; do
python tools/kbench.py $cfg 2>&1 | tail -1
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.log 2>&1; tail -1 gpurun_out/bench_s.log | cut -c1-260
