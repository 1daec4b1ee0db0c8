set -x
O=gpurun_out/r2
mkdir -p $O
nproc > $O/host.txt; free -g >> $O/host.txt; nvidia-smi --query-gpu=name,memory.total --format=csv >> $O/host.txt
python tools/r2_probe_cfg.py 24 5 128 1 65536,262144,1048576 3 > $O/probe_cfg4.log 2>&1
# per-launch DRAM bytes / L2 hit over one whole epoch (B=65536: launches 256..511)
ONLY_EPOCHS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none -k regex:force_batch --launch-skip 256 --launch-count 256 --csv --log-file $O/cfg4_dram_per_launch.csv python tools/r2_probe_cfg.py 24 5 128 1 65536 2 > $O/ncu_cfg4_a.log 2>&1
for skip in 256 384 500; do
ONLY_EPOCHS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_batch --launch-skip $skip --launch-count 1 -o /tmp/cfg4_full_l$skip -f python tools/r2_probe_cfg.py 24 5 128 1 65536 2 > $O/ncu_cfg4_b$skip.log 2>&1
python tools/ncu_summary.py /tmp/cfg4_full_l$skip.ncu-rep $O/cfg4_full_l$skip.md
done
ncu -i /tmp/cfg4_full_l384.ncu-rep --page source --csv > $O/cfg4_full_l384_source.csv 2>/dev/null
cp /tmp/cfg4_full_l384.ncu-rep $O/
# cfg3: RMAT-22 option 7 d=64 (force kernel + walk_kernel)
python tools/r2_probe_cfg.py 22 7 64 0 65536,262144 3 > $O/probe_cfg3.log 2>&1
ONLY_EPOCHS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"force_batch|walk_kernel" --launch-skip 65 --launch-count 3 -o /tmp/cfg3_full -f python tools/r2_probe_cfg.py 22 7 64 0 65536 2 > $O/ncu_cfg3.log 2>&1
python tools/ncu_summary.py /tmp/cfg3_full.ncu-rep $O/cfg3_full.md
du -sh $O
