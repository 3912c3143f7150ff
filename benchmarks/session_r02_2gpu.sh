# round-2 GPU session on 2 GPUs: tests, bench with the parity block, timeline, collective sweeps
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r02/pytest_gpu_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02/pytest_gpu_2gpu.log
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02/bench_n2.json 2> gpurun_out/r02/bench_n2.err; echo "bench exit $?" >> gpurun_out/r02/bench_n2.err
timeout 600 $TR --master-port 29512 benchmarks/step_timeline.py --out gpurun_out/r02/timeline_n2.json > gpurun_out/r02/timeline_n2.out 2> gpurun_out/r02/timeline_n2.err
timeout 600 $TR --master-port 29513 benchmarks/collective_sweep.py --out gpurun_out/r02/collective_sweep_n2.json > gpurun_out/r02/collective_sweep_n2.out 2>&1
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --microbench --max-log2 28 > gpurun_out/r02/microbench_n2.json 2> gpurun_out/r02/microbench_n2.err
timeout 300 python benchmarks/step_timeline.py --no-ddp --out gpurun_out/r02/timeline_n1b.json > /dev/null 2> gpurun_out/r02/timeline_n1b.err
tail -3 gpurun_out/r02/pytest_gpu_2gpu.log; tail -2 gpurun_out/r02/bench_n2.err
