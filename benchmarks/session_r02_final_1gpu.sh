mkdir -p gpurun_out/r02g
O=gpurun_out/r02g
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?" >> $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_arm.json 2> $O/bench_ref_arm.err; echo "ref exit $?" >> $O/bench_ref_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_n1.csv python bench.py --steps 1 --warmup 3 --no-ddp-leg --no-cpu-leg > $O/ncu_launchlist.log 2>&1
tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log; tail -1 $O/bench_n1.err; tail -1 $O/bench_ref_arm.err
