python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -3
for cfg in "conv 256 16 16 128 128 3" "conv 256 8 8 128 128 3" "conv 256 32 32 128 64 3" "conv 256 16 16 64 128 3" "conv 256 32 32 64 64 3" "conv 256 64 64 32 32 3" "conv 256 128 128 16 16 3"; do
python tools/kbench.py $cfg 2>&1 | tail -1
done
ncu --set full --clock-control none --import-source on -k regex:conv_tc_tma -c 1 -o gpurun_out/ncu_tma_64 python tools/kbench.py conv 256 32 32 64 64 3 1 > gpurun_out/ncu_tma_64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_stream -c 1 -o gpurun_out/ncu_stream_16b python tools/kbench.py conv 256 16 16 128 128 3 1 > gpurun_out/ncu_stream_16b.log 2>&1
