timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_multi_rank.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 100 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print(round(d['ms_per_step'],4), {a:round(b,4) for a,b in k.items() if b>0}, d['parity']['ok'])"
