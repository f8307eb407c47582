# prologue / fused-upsample variants of the resident-filter conv against the plain TMA kernel (decides what the model path uses)
for shape in "256 128 128 16 16 3" "256 64 64 32 32 3" "256 32 32 64 64 3" "256 128 128 16 16 1" "256 64 64 32 16 3"; do
  python tools/kbench.py conv $shape 2>&1 | tail -1 | sed "s/^/[plain] /"
  KB_PRE=1 python tools/kbench.py conv $shape 2>&1 | tail -1 | sed "s/^/[pre] /"
  KB_PRE=0 python tools/kbench.py conv $shape 2>&1 | tail -1 | sed "s/^/[pre-entry, no prologue] /"
  KB_PRE=1 TTG_ROWS=1 python tools/kbench.py conv $shape 2>&1 | tail -1 | sed "s/^/[pre rows] /"
done
for shape in "256 128 128 32 16 3" "256 64 64 64 32 3" "256 32 32 128 64 3" "256 128 128 16 16 3"; do
  KB_PRE=1 KB_UP=1 python tools/kbench.py conv $shape 2>&1 | tail -1 | sed "s/^/[pre up] /"
  KB_PRE=0 KB_UP=1 python tools/kbench.py conv $shape 2>&1 | tail -1 | sed "s/^/[up] /"
done
python tools/kbench.py bn 4194304 16 2>&1 | tail -3
python tools/kbench.py bn 1048576 32 2>&1 | tail -3
python tools/kbench.py bn 262144 64 2>&1 | tail -3
