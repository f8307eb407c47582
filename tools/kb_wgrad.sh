for m in 0 1; do
 for shape in "256 128 128 16 16 3" "256 64 64 32 32 3" "256 64 64 16 32 3" "256 128 128 32 16 3" "256 64 64 64 32 3" "256 32 32 64 64 3"; do
  TTG_MFOLD=$m python tools/kbench.py wgrad $shape 2>&1 | tail -1 | sed "s/^/mfold=$m /"
 done
done
