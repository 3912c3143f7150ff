# round-2 8-GPU session B: the pipelined NVLS reduce at N=8 (correctness, then config-5 numbers)
mkdir -p gpurun_out/r02f
O=gpurun_out/r02f
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B200DP_TIMEOUT_MS=30000 B200DP_STAGING_MB=32 timeout 200 $TR --nproc-per-node 8 --master-port 29531 tests/scripts/dist_collectives_large.py > $O/collectives_large_n8.log 2>&1; RC=$?; echo "exit $RC" >> $O/collectives_large_n8.log
if [ $RC -ne 0 ]; then export B200DP_PIPE_REDUCE=0; echo "PIPELINED REDUCE DISABLED FOR THE REST OF THE SESSION" >> $O/collectives_large_n8.log; fi
timeout 300 $TR --nproc-per-node 8 --master-port 29537 benchmarks/collective_sweep.py --blocks 148 --sizes-mib 16,64,256 --out $O/collective_sweep_n8.json > $O/collective_sweep_n8.out 2>&1
timeout 500 $TR --nproc-per-node 8 --master-port 29536 bench.py --gpus 8 --microbench --max-log2 30 > $O/microbench_n8.json 2> $O/microbench_n8.err
tail -2 $O/collectives_large_n8.log | cut -c1-200; grep MiB $O/collective_sweep_n8.out | cut -c1-400
