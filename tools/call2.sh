set -x
O=gpurun_out/r02s
mkdir -p $O
for f in 0 32 64 128; do tools/bin/random_sectors 2048 64 0 $f; done > $O/random_sectors_fetch.log 2>&1
for f in 32 128; do tools/bin/random_sectors 2048 64 1 $f; done >> $O/random_sectors_fetch.log 2>&1
timeout 600 python tools/sweep.py kmer=0 kmer=0,l2_fetch_parse=32 kmer=0,l2_fetch_parse=0,l2_fetch=32 kmer=0,l2_fetch=128 kmer=0,l2_fetch=64 kmer=0,l2_fetch=32 > $O/sweep_l2fetch.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "suffix_array or kmer_table or golden or sort" > $O/pytest_index.log 2>&1; echo "pytest rc $?" >> $O/pytest_index.log
cat $O/random_sectors_fetch.log; cat $O/sweep_l2fetch.log | cut -c1-400; tail -3 $O/pytest_index.log
