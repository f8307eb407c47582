for nb in 0 1; do for ps in 1 2 3 4 6 8; do
 for shape in "256 128 128 16 16 3" "256 64 64 32 32 3" "256 128 128 16 16 1"; do
  env TTG_NBUF=$nb TTG_PERSM=$ps python tools/kbench.py conv $shape 2>&1 | tail -1 | sed "s/^/[nbuf=$nb persm=$ps] /"
 done
done; done
