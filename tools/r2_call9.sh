set -x
O=gpurun_out/r2i
mkdir -p $O
for c in 2048 8192; do CHUNK=$c ONLY_EPOCHS=1 python tools/r2_probe_cfg.py 24 5 128 1 262144 3 2>&1 | tail -1 | cut -c1-200 | tee -a $O/chunk_cfg4.log; done
for c in 128 512 1024; do CHUNK=$c ONLY_EPOCHS=1 python tools/r2_probe_cfg.py 24 5 128 1 65536 3 2>&1 | tail -1 | cut -c1-200 | tee -a $O/chunk_cfg4_B65536.log; done
for c in 128 256 512 1024; do CHUNK=$c ONLY_EPOCHS=1 python tools/r2_probe_cfg.py 20 6 128 0 65536,16384,4096 4 2>&1 | tail -3 | cut -c1-200 | tee -a $O/chunk_cfg2.log; done
for c in 128 1024; do CHUNK=$c ONLY_EPOCHS=1 python tools/r2_probe_cfg.py 22 6 64 0 65536 4 2>&1 | tail -1 | cut -c1-200 | tee -a $O/chunk_rmat22_opt6_d64.log; done
