set -x
python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/prof_r1_final.json > gpurun_out/bench_final_1gpu.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_tma -c 1 -o gpurun_out/ncu_f_conv16 python tools/kbench.py conv 256 128 128 16 16 3 1 > gpurun_out/ncu_f_conv16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_stream -c 1 -o gpurun_out/ncu_f_stream128 python tools/kbench.py conv 256 16 16 128 128 3 1 > gpurun_out/ncu_f_stream128.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chan_ -s 30 -c 2 -o gpurun_out/ncu_f_bn_bwd python tools/kbench.py bn 4194304 16 > gpurun_out/ncu_f_bn_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -c 1 -o gpurun_out/ncu_f_wgrad64 python tools/kbench.py wgrad 256 32 32 64 64 3 1 > gpurun_out/ncu_f_wgrad64.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --eager --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1
# --- session 3 additions
python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 2 -c 1 -o gpurun_out/ncu_attn_fwd python tools/attn_one.py 16 4096 1024 8 32 > gpurun_out/ncu_attn_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 2 -c 1 -o gpurun_out/ncu_attn_bwd python tools/attn_one.py 16 4096 1024 8 32 > gpurun_out/ncu_attn_bwd.log 2>&1
python tools/bench_configs.py --steps 10 --warmup 4 > gpurun_out/bench_configs.log 2>&1
# 8 GPUs: python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/bench_configs.py --fused-only
# run-to-run spread of the fp32 oracle-width test: python tools/flaky_probe.py 16
# A/B of the launch-count changes: TTG_NO_DIRECT=1 / TTG_NO_ARENA=1 python bench.py --no-cpu-baseline
# e2e with the CPU draws overlapping the generator-sample graph (unmeasured, off by default): TTG_EARLY_GEN=1 python bench.py --no-cpu-baseline
