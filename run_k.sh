set -x
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-200 gpurun_out/r2_bench_n1.json
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_n200.csv python bench.py --size 200 --steps 8 --warmup 3 --no-cpu --no-parity > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_geom_tiles_f|k_predict|k_edge_constraints|k_commit" --launch-skip 12 --launch-count 4 -f -o gpurun_out/r2_final_n200 python bench.py --size 200 --steps 4 --warmup 3 --no-cpu --no-parity > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/r2_final_n200.ncu-rep
