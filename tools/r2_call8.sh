set -x
O=gpurun_out/r2h
mkdir -p $O
(CUDA_VISIBLE_DEVICES=1 CHUNK=128 python tools/r2_probe_cfg.py 24 5 128 1 262144 3 > $O/n1_chunk128.log 2>&1; CUDA_VISIBLE_DEVICES=1 CHUNK=256 python tools/r2_probe_cfg.py 24 5 128 1 262144 3 > $O/n1_chunk256.log 2>&1; CUDA_VISIBLE_DEVICES=1 CHUNK=512 python tools/r2_probe_cfg.py 24 5 128 1 262144 3 > $O/n1_chunk512.log 2>&1; CUDA_VISIBLE_DEVICES=1 CHUNK=1024 python tools/r2_probe_cfg.py 24 5 128 1 262144 3 > $O/n1_chunk1024.log 2>&1; tail -q -n1 $O/n1_chunk*.log | cut -c1-330)
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -k "one_process or failure" > $O/pytest_inproc.log 2>&1; tail -3 $O/pytest_inproc.log
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
for MC in 1 0; do
MC=$MC SCALE=24 MODEL=5 DIM=128 BS=1 BATCHES=262144 CHUNKS=128,256 ORDERS=1 SIGS=2 FREE=0 TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 tools/mgpu_probe.py > $O/mgpu_cfg4_n2_mc$MC.log 2> $O/mgpu_cfg4_n2_mc$MC.err
tail -2 $O/mgpu_cfg4_n2_mc$MC.err; cut -c1-330 $O/mgpu_cfg4_n2_mc$MC.log
done
