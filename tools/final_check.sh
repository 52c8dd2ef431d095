# the round-end sequence on one B200: GPU tests, smoke, both bench arms, and the ncu launch list of the bench command
set -x
O=gpurun_out/${1:-final3}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/smoke.log
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc $?" >> $O/bench_1gpu.err
[ "${2:-ncu}" = noncu ] || timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/ncu_bench.log 2>&1
tail -3 $O/pytest_gpu.log $O/smoke.log; tail -2 $O/bench_1gpu.err; wc -c $O/*
