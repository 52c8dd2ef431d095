set -x
O=gpurun_out/r02w
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmer_table or suffix_array or golden or parse_random or alphabet or dense" > $O/pytest_table.log 2>&1; echo "pytest rc $?" >> $O/pytest_table.log
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "n_runs or index_moves" >> $O/pytest_table.log 2>&1; echo "pytest2 rc $?" >> $O/pytest_table.log
timeout 200 python tools/index_probe.py > $O/index_probe.log 2>&1
tail -4 $O/pytest_table.log; cut -c1-200 $O/index_probe.log
