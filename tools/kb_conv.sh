# usage: tools/kb_conv.sh "ENV=VAL ..." ["ENV=VAL ..." ...]   one pass over the headline conv shapes per environment
for envs in "$@"; do
 for shape in "256 128 128 16 16 3" "256 64 64 32 32 3" "256 32 32 64 64 3" "256 128 128 16 16 1" "256 64 64 32 16 3" "256 64 64 16 32 3"; do
  env $envs python tools/kbench.py conv $shape 2>&1 | tail -1 | sed "s/^/[$envs] /"
 done
done
