set -x
O=gpurun_out/r2l
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_cfg4_n1.json 2> $O/bench_cfg4_n1.err; tail -2 $O/bench_cfg4_n1.err; cut -c1-3500 $O/bench_cfg4_n1.json
# per-launch DRAM bytes for the default workload (B=262144: launches 64..127 = the 2nd epoch)
ONLY_EPOCHS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:force_batch --launch-skip 64 --launch-count 64 --csv --log-file $O/cfg4_dram_per_launch_B262144.csv python tools/r2_probe_cfg.py 24 5 128 1 262144 2 > $O/ncu_cfg4.log 2>&1; tail -1 $O/ncu_cfg4.log | cut -c1-200
