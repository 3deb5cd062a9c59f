(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_multi_rank.py tests/test_gpu_layers.py tests/test_gpu_boundary.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5)
for v in default pt2 et3 nopt; do
  unset SMGPU_LIB SMGPU_NO_POINT_TILES
  if [ $v = pt2 ] || [ $v = et3 ]; then export SMGPU_LIB=$PWD/variants/libsmgpu_$v.so; fi
  if [ $v = nopt ]; then export SMGPU_NO_POINT_TILES=1; fi
  timeout 600 python bench.py --size 200 --steps 20 --warmup 3 --no-cpu --no-parity > gpurun_out/r2_b200_e_$v.json 2> gpurun_out/r2_b200_e_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_b200_e_$v.json')); print('$v', round(d['ms_per_step'],4), d['config']['setup_s']['create_upload'], d['config']['hbm_resident_gb'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0})"
done
