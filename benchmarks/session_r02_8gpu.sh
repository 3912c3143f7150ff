# round-2 GPU session on 8 GPUs: correctness of the NVLS / pipelined kernels at N=8, headline bench, timeline, config 4 + 5
mkdir -p gpurun_out/r02e
O=gpurun_out/r02e
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 600 -k "large_gather" > $O/pytest_large.log 2>&1; echo "pytest exit $?" >> $O/pytest_large.log
B200DP_TIMEOUT_MS=30000 B200DP_STAGING_MB=32 timeout 200 $TR --nproc-per-node 8 --master-port 29531 tests/scripts/dist_collectives_large.py > $O/collectives_large_n8.log 2>&1; RC=$?; echo "exit $RC" >> $O/collectives_large_n8.log
if [ $RC -ne 0 ]; then export B200DP_PIPE_REDUCE=0; echo "PIPELINED REDUCE DISABLED FOR THE REST OF THE SESSION" >> $O/collectives_large_n8.log; fi
B200DP_TIMEOUT_MS=30000 timeout 200 $TR --nproc-per-node 8 --master-port 29532 tests/scripts/dist_allreduce_check.py > $O/allreduce_check_n8.log 2>&1; echo "exit $?" >> $O/allreduce_check_n8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench exit $?" >> $O/bench_n8.err
B200DP_NVLS_BLOCKS=8 timeout 400 $TR --nproc-per-node 8 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 --no-ddp-leg --no-parity > $O/bench_n8_nvls8.json 2> $O/bench_n8_nvls8.err
timeout 400 $TR --nproc-per-node 8 --master-port 29535 benchmarks/step_timeline.py --no-ddp --out $O/timeline_n8.json > $O/timeline_n8.out 2> $O/timeline_n8.err
timeout 500 $TR --nproc-per-node 8 --master-port 29536 bench.py --gpus 8 --microbench --max-log2 30 > $O/microbench_n8.json 2> $O/microbench_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29537 benchmarks/collective_sweep.py --blocks 64,148 --sizes-mib 16,64,256 --out $O/collective_sweep_n8.json > $O/collective_sweep_n8.out 2>&1
timeout 600 $TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --model bert-large --accum 4 --steps 6 --warmup 3 > $O/bench_bert_n8.json 2> $O/bench_bert_n8.err
tail -2 $O/pytest_large.log; tail -2 $O/collectives_large_n8.log; tail -1 $O/allreduce_check_n8.log; tail -1 $O/bench_n8.err
