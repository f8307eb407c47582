ncu --set full --clock-control none --import-source on -k regex:chan_reduce -c 2 -o gpurun_out/ncu_bn_reduce python tools/kbench.py bn 4194304 16 > gpurun_out/ncu_bn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chan_map -c 2 -o gpurun_out/ncu_bn_map python tools/kbench.py bn 4194304 16 >> gpurun_out/ncu_bn.log 2>&1
