ncu --set full --clock-control none --import-source on -k regex:chan_ -s 30 -c 2 -o gpurun_out/ncu_bn_bwd python tools/kbench.py bn 4194304 16 > gpurun_out/ncu_bn_bwd.log 2>&1
