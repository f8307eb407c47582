for t in "4,4" "8,4" "8,3" "8,2" "6,4" "6,2" "8,1"; do
 for shape in "256 128 128 16 16 3" "256 64 64 32 32 3" "256 64 64 64 32 3"; do
  TTG_WG_TUNE=$t python tools/kbench.py wgrad $shape 2>&1 | tail -1 | sed "s/^/[nbuf,persm=$t] /"
 done
done
