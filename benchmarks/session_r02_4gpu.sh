# round-2 GPU session on 4 GPUs: NVLS parity + N=4 bench / timeline / collective sweeps
mkdir -p gpurun_out/r02c
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_accelerator.py -m gpu -q --timeout 900 -k "four_ranks or large_gather or three_ranks or deferred or fp16 or resume or join or unused" > gpurun_out/r02c/pytest_4gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c/pytest_4gpu.log
B200DP_TIMEOUT_MS=60000 timeout 300 $TR --nproc-per-node 4 --master-port 29520 tests/scripts/dist_allreduce_check.py > gpurun_out/r02c/allreduce_check_n4.log 2>&1; echo "exit $?" >> gpurun_out/r02c/allreduce_check_n4.log
timeout 600 $TR --nproc-per-node 4 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02c/bench_n4.json 2> gpurun_out/r02c/bench_n4.err; echo "bench exit $?" >> gpurun_out/r02c/bench_n4.err
timeout 600 $TR --nproc-per-node 4 --master-port 29522 benchmarks/step_timeline.py --no-ddp --out gpurun_out/r02c/timeline_n4.json > gpurun_out/r02c/timeline_n4.out 2> gpurun_out/r02c/timeline_n4.err
timeout 600 $TR --nproc-per-node 4 --master-port 29523 benchmarks/collective_sweep.py --blocks 16,32,64,148 --out gpurun_out/r02c/collective_sweep_n4.json > gpurun_out/r02c/collective_sweep_n4.out 2>&1
timeout 600 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --microbench --max-log2 28 > gpurun_out/r02c/microbench_n4.json 2> gpurun_out/r02c/microbench_n4.err
tail -3 gpurun_out/r02c/pytest_4gpu.log; tail -2 gpurun_out/r02c/bench_n4.err; tail -2 gpurun_out/r02c/allreduce_check_n4.log
