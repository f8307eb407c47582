# Round-2 evidence (one gpurun call, one GPU): every ncu command is preceded by the same command without ncu.
# The .ncu-rep files (40 MB each with --import-source) are summarised ON the box (tools/ncu_hot.py) and deleted, so that
# gpurun_out/ stays under the 64 MiB that travel back.
# usage: bash tools/evidence_round2.sh   -> gpurun_out/r2_ncu_*.txt, gpurun_out/r2_launches.csv
set -x
cap() {   # cap <tag> <kernel regex> <skip> <count> <cmd...>
  tag=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > gpurun_out/r2_plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/r2_ncu_$tag "$@" > gpurun_out/r2_ncu_$tag.log 2>&1
  echo "$tag rc=$?"
  { echo "# $* (ncu --set full --clock-control none, launch $skip of kernel /$rx/)"; cat gpurun_out/r2_plain_$tag.log;
    python tools/ncu_hot.py gpurun_out/r2_ncu_$tag.ncu-rep 14; } > gpurun_out/r2_ncu_$tag.txt 2>&1
  rm -f gpurun_out/r2_ncu_$tag.ncu-rep
}
cap conv16   conv_tc_tma        3 1 python tools/kbench.py conv 256 128 128 16 16 3 2
cap conv32   conv_tc_tma        3 1 python tools/kbench.py conv 256 64 64 32 32 3 2
cap conv64   conv_tc_tma        3 1 python tools/kbench.py conv 256 32 32 64 64 3 2
cap stream128 conv_tc_stream    3 1 python tools/kbench.py conv 256 16 16 128 128 3 2
cap wgrad16  conv_wgrad         3 1 python tools/kbench.py wgrad 256 128 128 16 16 3 2
cap wgrad64  conv_wgrad         3 1 python tools/kbench.py wgrad 256 32 32 64 64 3 2
cap bn16     chan_              30 2 python tools/kbench.py bn 4194304 16
python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_plain_bench_eager.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv --log-file gpurun_out/r2_launches.csv python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
du -sh gpurun_out
