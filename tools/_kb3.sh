for mode in 1 2 3; do
echo "=== TTG_TMA_MODE=$mode"
export TTG_TMA_MODE=$mode
for cfg in "conv 256 128 128 16 16 3" "conv 256 64 64 32 32 3" "conv 256 32 32 64 64 3" "conv 256 64 64 16 32 3" "conv 256 32 32 64 32 3" "conv 256 64 64 32 32 1"; do
python tools/kbench.py $cfg 2>&1 | tail -1
done
done
