mkdir -p gpurun_out/r02i
O=gpurun_out/r02i
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-leg > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?" >> $O/bench_n1.err
tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log; tail -1 $O/bench_n1.err
