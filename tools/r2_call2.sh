set -x
O=gpurun_out/r2b
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
python tools/r2_tune_ring.py 24 5 128 1 65536 -1,20,21,22,23,-1 > $O/tune_cfg4.log 2>&1; cat $O/tune_cfg4.log | cut -c1-330
python tools/r2_tune_ring.py 20 6 128 0 65536 -1,20,21,22,23 > $O/tune_cfg2.log 2>&1; cat $O/tune_cfg2.log | cut -c1-330
python tools/r2_tune_ring.py 22 7 64 0 65536 -1,20,21,22 > $O/tune_cfg3.log 2>&1; cat $O/tune_cfg3.log | cut -c1-330
python tools/r2_tune_ring.py 20 6 128 0 256 -1,20,21,22,11 > $O/tune_cfg2_B256.log 2>&1; cat $O/tune_cfg2_B256.log | cut -c1-330
python tools/r2_tune_ring.py 20 5 128 1 4096 -1,20,21,22,11 > $O/tune_rmat20_opt5_B4096.log 2>&1; cat $O/tune_rmat20_opt5_B4096.log | cut -c1-330
python bench.py --workload cfg2 --steps 2 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; tail -3 $O/bench_cfg2.err; cut -c1-1500 $O/bench_cfg2.json
python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err; tail -3 $O/bench_cfg4.err; cut -c1-3000 $O/bench_cfg4.json
