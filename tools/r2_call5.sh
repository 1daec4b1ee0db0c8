set -x
O=gpurun_out/r2e
mkdir -p $O
for g in 8 16 24 32 48; do ./tools/probes/random_gather_probe $g 512 2>&1 | head -3 | tee -a $O/random_gather_sizes.log; done
VARIANT=22 ONLY_EPOCHS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_batch --launch-skip 384 --launch-count 1 -o /tmp/cfg4_ring22_l384 -f python tools/r2_probe_cfg.py 24 5 128 1 65536 2 > $O/ncu_ring.log 2>&1
python tools/ncu_summary.py /tmp/cfg4_ring22_l384.ncu-rep $O/cfg4_ring22_l384.md
ncu -i /tmp/cfg4_ring22_l384.ncu-rep --page details > $O/cfg4_ring22_l384_details.txt 2>/dev/null
ncu -i /tmp/cfg4_ring22_l384.ncu-rep --page source --csv > $O/cfg4_ring22_l384_source.csv 2>/dev/null
VARIANT=-1 ONLY_EPOCHS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_batch --launch-skip 384 --launch-count 1 -o /tmp/cfg4_reg_l384 -f python tools/r2_probe_cfg.py 24 5 128 1 65536 2 > $O/ncu_reg.log 2>&1
python tools/ncu_summary.py /tmp/cfg4_reg_l384.ncu-rep $O/cfg4_reg_clampless_l384.md
ncu -i /tmp/cfg4_reg_l384.ncu-rep --page details > $O/cfg4_reg_clampless_l384_details.txt 2>/dev/null
du -sh $O
