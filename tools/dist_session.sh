timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -s > gpurun_out/dist_tests.log 2>&1; echo "dist tests rc=$?"; grep "parity\|passed\|failed\|shutdown" gpurun_out/dist_tests.log | tail -12
for blocking in 0 1; do
TTG_DP_BLOCKING=$blocking timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_2gpu_blocking$blocking.log 2>&1; echo "bench2 blocking=$blocking rc=$?"; grep '"metric"' gpurun_out/bench_2gpu_blocking$blocking.log | cut -c1-330
done
