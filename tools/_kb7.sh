python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
for cfg in "conv 256 16 16 128 128 3" "conv 256 8 8 128 128 3" "conv 256 16 16 128 64 3" "conv 256 16 16 64 128 3" "conv 256 32 32 128 64 3"; do
python tools/kbench.py $cfg 2>&1 | tail -1
done
