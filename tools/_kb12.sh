python tools/kbench.py bn 4194304 16 | tail -3; python tools/kbench.py bn 1048576 32 | tail -3; python tools/kbench.py bn 65536 128 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_b.log 2>&1
tail -2 gpurun_out/ncu_bench_b.log | cut -c1-200
