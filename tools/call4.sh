set -x
O=gpurun_out/r02u
mkdir -p $O
timeout 400 python tools/sweep.py kmer=0 kmer=0,spec_sync=2 kmer=0,spec_sync=2,occ=6 kmer=0,spec_sync=2,occ=10 kmer=0,spec_sync=2,occ=12 kmer=0,spec_sync=1,occ=8 > $O/sweep_spec_queue.log 2>&1
SWEEP_P_SNP=8e-3 SWEEP_P_INS=1e-3 SWEEP_P_DEL=1e-3 timeout 400 python tools/sweep.py kmer=0 kmer=0,spec_sync=2 kmer=0,spec_sync=2,occ=6 kmer=0,spec_sync=1,occ=8 > $O/sweep_spec_queue_cfg5rates.log 2>&1
SWEEP_P_SNP=1e-3 SWEEP_P_INS=1e-4 SWEEP_P_DEL=1e-4 timeout 400 python tools/sweep.py kmer=0 kmer=0,spec_sync=2 kmer=0,spec_sync=1,occ=8 > $O/sweep_spec_queue_indel1e-4.log 2>&1
cut -c1-330 $O/sweep_spec_queue.log $O/sweep_spec_queue_cfg5rates.log $O/sweep_spec_queue_indel1e-4.log
