set -x
python tools/kbench.py conv 256 16 16 128 128 3 2>&1 | tail -1
python tools/kbench.py conv 256 8 8 128 128 3 2>&1 | tail -1
python tools/kbench.py conv 256 32 32 64 64 3 2>&1 | tail -1
python tools/kbench.py conv 256 64 64 32 32 3 2>&1 | tail -1
python tools/kbench.py conv 256 128 128 16 16 3 2>&1 | tail -1
python tools/kbench.py conv 256 16 16 64 128 3 2>&1 | tail -1
python tools/kbench.py wgrad 256 16 16 128 128 3 2>&1 | tail -1
python tools/kbench.py wgrad 256 8 8 128 128 3 2>&1 | tail -1
python tools/kbench.py wgrad 256 32 32 64 64 3 2>&1 | tail -1
python tools/kbench.py wgrad 256 128 128 16 16 3 2>&1 | tail -1
python tools/kbench.py bn 4194304 16 2>&1 | tail -3
python tools/kbench.py bn 65536 128 2>&1 | tail -3
python tools/kbench.py bn 16384 128 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:conv_tc_stream -c 1 -o gpurun_out/ncu_stream_16 python tools/kbench.py conv 256 16 16 128 128 3 1 > gpurun_out/ncu_stream_16.log 2>&1
