mkdir -p gpurun_out/r02d
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 900 -k "large_gather" > gpurun_out/r02d/pytest_large.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d/pytest_large.log
timeout 400 $TR --nproc-per-node 4 --master-port 29523 benchmarks/collective_sweep.py --blocks 148 --sizes-mib 16,64,256 --out gpurun_out/r02d/collective_sweep_n4.json > gpurun_out/r02d/collective_sweep_n4.out 2>&1
B200DP_PIPE_CHUNK_MB=4 timeout 400 $TR --nproc-per-node 4 --master-port 29525 benchmarks/collective_sweep.py --blocks 148 --sizes-mib 64,256 > gpurun_out/r02d/collective_sweep_n4_chunk4.out 2>&1
B200DP_PIPE_CHUNK_MB=16 timeout 400 $TR --nproc-per-node 4 --master-port 29526 benchmarks/collective_sweep.py --blocks 148 --sizes-mib 64,256 > gpurun_out/r02d/collective_sweep_n4_chunk16.out 2>&1
timeout 400 $TR --nproc-per-node 2 --master-port 29524 benchmarks/collective_sweep.py --blocks 148 --sizes-mib 16,64,256 --out gpurun_out/r02d/collective_sweep_n2.json > gpurun_out/r02d/collective_sweep_n2.out 2>&1
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02d/smoke.log 2>&1
tail -3 gpurun_out/r02d/pytest_large.log; tail -4 gpurun_out/r02d/collective_sweep_n4.out | cut -c1-600
