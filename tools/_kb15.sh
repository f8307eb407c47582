timeout 300 python -m pytest tests/test_gpu_paths.py -x -q -m gpu -k "folded" 2>&1 | tail -5
for f in 1 0; do
for cfg in "conv 256 128 128 16 16 3" "conv 256 64 64 32 32 3"; do
TTG_FOLD=$f timeout 120 python tools/kbench.py $cfg 2>&1 | tail -1
done
done
TTG_B200_LIB=tartangan_b200/lib/libttg_b200_trace.so python tools/trace_stream.py 256 128 128 16 16 3 > gpurun_out/trace_fold2.log 2>&1
