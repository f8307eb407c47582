for lib in "" "tartangan_b200/lib/libttg_b200_hint.so"; do
echo "=== lib=$lib"
export TTG_B200_LIB=$lib
[ -z "$lib" ] && unset TTG_B200_LIB
for cfg in "conv 256 16 16 128 128 3" "conv 256 8 8 128 128 3" "conv 256 32 32 64 64 3" "conv 256 64 64 32 32 3" "conv 256 128 128 16 16 3" "conv 256 16 16 64 128 3" "wgrad 256 16 16 128 128 3" "wgrad 256 8 8 128 128 3" "wgrad 256 32 32 64 64 3" "wgrad 256 128 128 16 16 3" "wgrad 256 64 64 32 32 3"; do
python tools/kbench.py $cfg 2>&1 | tail -1
done
done
