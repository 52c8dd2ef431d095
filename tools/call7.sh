set -x
O=gpurun_out/r02v
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -x -q > $O/pytest_cli.log 2>&1; echo "pytest rc $?" >> $O/pytest_cli.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_sort_scatter2|k_sa_finish0|k_tab_mark" --launch-skip 9 -c 4 -o $O/index_kernels_final python tools/index_probe.py > $O/ncu_index.log 2>&1
tail -5 $O/pytest_cli.log
