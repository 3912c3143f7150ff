# round-2 GPU session B on 2 GPUs
mkdir -p gpurun_out/r02b
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r02b/pytest_gpu_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b/pytest_gpu_2gpu.log
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b/bench_n2.json 2> gpurun_out/r02b/bench_n2.err; echo "bench exit $?" >> gpurun_out/r02b/bench_n2.err
B200DP_TWOSHOT_BLOCKS=16 timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --no-ddp-leg --no-parity > gpurun_out/r02b/bench_n2_ctas16.json 2> gpurun_out/r02b/bench_n2_ctas16.err
B200DP_FAST_HOOKS=0 timeout 600 $TR --master-port 29516 bench.py --gpus 2 --steps 20 --warmup 5 --no-ddp-leg --no-parity > gpurun_out/r02b/bench_n2_slowhooks.json 2> gpurun_out/r02b/bench_n2_slowhooks.err
timeout 600 $TR --master-port 29512 benchmarks/step_timeline.py --no-ddp --out gpurun_out/r02b/timeline_n2.json > gpurun_out/r02b/timeline_n2.out 2> gpurun_out/r02b/timeline_n2.err
timeout 300 $TR --master-port 29513 benchmarks/pack_dst_probe.py > gpurun_out/r02b/pack_dst_probe_n2.json 2> gpurun_out/r02b/pack_dst_probe_n2.err
timeout 300 python benchmarks/pack_dst_probe.py > gpurun_out/r02b/pack_dst_probe_n1.json 2> gpurun_out/r02b/pack_dst_probe_n1.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-ddp-leg --no-cpu-leg > gpurun_out/r02b/bench_n1.json 2> gpurun_out/r02b/bench_n1.err
timeout 600 $TR --master-port 29517 bench.py --gpus 2 --model bert-large --accum 4 --steps 6 --warmup 3 > gpurun_out/r02b/bench_bert_n2.json 2> gpurun_out/r02b/bench_bert_n2.err
tail -3 gpurun_out/r02b/pytest_gpu_2gpu.log; tail -2 gpurun_out/r02b/bench_n2.err
