SMGPU_TIMING=1 timeout 600 python bench.py 2>gpurun_out/bench_setup.err | tee gpurun_out/bench_setup.json | cut -c1-200
grep -a "smgpu create" gpurun_out/bench_setup.err | tail -12
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
