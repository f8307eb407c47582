python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
for cfg in "wgrad 256 16 16 128 128 3" "wgrad 256 8 8 128 128 3" "wgrad 256 32 32 64 64 3" "wgrad 256 64 64 32 32 3" "wgrad 256 128 128 16 16 3" "wgrad 256 16 16 64 128 3" "wgrad 256 128 128 32 16 3" "wgrad 256 32 32 16 32 1"; do
python tools/kbench.py $cfg 2>&1 | tail -1
done
