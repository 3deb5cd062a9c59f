timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --n 128 --min-angle 80 --max-angle 100 --steps 20 --no-cpu 2>/dev/null | tee gpurun_out/r2_dense128_b.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dense:', d['ms_per_step'], d['kernel_ms_per_step'], d['parity'])"
