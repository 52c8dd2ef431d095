set -x
O=gpurun_out/r02t
mkdir -p $O
SORT_VARIANTS=0,1,2,0 timeout 200 python tools/index_probe.py > $O/index_probe_sort_variants.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "suffix_array or kmer_table or radix_sort" > $O/pytest_index.log 2>&1; echo "pytest rc $?" >> $O/pytest_index.log
timeout 900 python tools/big_check.py > $O/big_check_3100M.log 2>&1
cat $O/index_probe_sort_variants.log | cut -c1-330; tail -3 $O/pytest_index.log; grep -n "gpu index\|PASS\|FAIL" $O/big_check_3100M.log | cut -c1-400
