mkdir -p gpurun_out/r02j
O=gpurun_out/r02j
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 600 -k "large_gather and not 4-8" > $O/pytest_large.log 2>&1; echo "pytest exit $?" >> $O/pytest_large.log
timeout 300 $TR --nproc-per-node 2 --master-port 29524 benchmarks/collective_sweep.py --blocks 148 --sizes-mib 16,64,256 --out $O/collective_sweep_n2.json > $O/collective_sweep_n2.out 2>&1
B200DP_PUSH2_ROUND_KB=512 timeout 300 $TR --nproc-per-node 2 --master-port 29525 benchmarks/collective_sweep.py --blocks 148 --sizes-mib 64,256 > $O/collective_sweep_n2_round512.out 2>&1
timeout 300 $TR --nproc-per-node 2 --master-port 29526 bench.py --gpus 2 --microbench --max-log2 28 > $O/microbench_n2.json 2> $O/microbench_n2.err
tail -3 $O/pytest_large.log; grep MiB $O/collective_sweep_n2.out | cut -c1-500
