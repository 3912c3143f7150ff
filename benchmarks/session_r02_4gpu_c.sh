mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B200DP_TIMEOUT_MS=30000 B200DP_STAGING_MB=32 timeout 200 $TR --nproc-per-node 4 --master-port 29531 tests/scripts/dist_collectives_large.py > $O/collectives_large_n4.log 2>&1; echo "exit $?" >> $O/collectives_large_n4.log
timeout 400 $TR --nproc-per-node 4 --master-port 29536 bench.py --gpus 4 --microbench --max-log2 28 > $O/microbench_n4.json 2> $O/microbench_n4.err
timeout 400 $TR --nproc-per-node 4 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 --no-ddp-leg > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench exit $?" >> $O/bench_n4.err
tail -1 $O/collectives_large_n4.log; tail -1 $O/bench_n4.err
