set -x
mkdir -p gpurun_out/r02r
O=gpurun_out/r02r
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "suffix_array or kmer_table or golden or parse_random or sort" > $O/pytest_index.log 2>&1; echo "pytest rc $?" >> $O/pytest_index.log
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "index_moves or config3_shape" >> $O/pytest_index.log 2>&1; echo "pytest2 rc $?" >> $O/pytest_index.log
timeout 120 python tools/index_probe.py > $O/index_probe.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_sort_scatter|k_sa_finish0|k_sort_hist" --launch-skip 8 -c 5 -o $O/index_kernels python tools/index_probe.py > $O/ncu_index.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_index.csv python tools/index_probe.py > $O/ncu_launches.log 2>&1
timeout 300 python tools/scale_check.py > $O/scale_check_249M.log 2>&1
timeout 600 python tools/big_check.py > $O/big_check_3100M.log 2>&1
tail -3 $O/pytest_index.log $O/index_probe.log $O/scale_check_249M.log $O/big_check_3100M.log
