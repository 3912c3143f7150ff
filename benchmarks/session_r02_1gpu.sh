mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02/gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r02/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err; echo "bench exit $?" >> gpurun_out/r02/bench_n1.err
timeout 600 python benchmarks/pack_sweep.py > gpurun_out/r02/pack_sweep.jsonl 2>&1
timeout 600 python benchmarks/step_timeline.py --out gpurun_out/r02/timeline_n1.json > gpurun_out/r02/timeline_n1.out 2> gpurun_out/r02/timeline_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:multi_adam -s 1 -c 1 -f -o gpurun_out/r02/multi_adam python benchmarks/kernel_probe.py --adam > gpurun_out/r02/ncu_adam.log 2>&1
python - > gpurun_out/r02/nvml_probe.txt 2>&1 <<'PY'
import pynvml, torch
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for fid in (138, 139, 140, 141):
    for scope in (0xFFFFFFFF, 0):
        try:
            v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, scope)])[0]
            print(fid, scope, v.nvmlReturn, v.valueType, v.value.ullVal)
        except Exception as e:
            print(fid, scope, "ERR", e)
import subprocess
print(subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", "0"], capture_output=True, text=True).stdout[:1500])
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
PY
tail -3 gpurun_out/r02/pytest_gpu.log; cat gpurun_out/r02/bench_n1.err | tail -3
