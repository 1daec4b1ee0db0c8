set -x
O=gpurun_out/r2g
mkdir -p $O
SCALE=24 MODEL=5 DIM=128 BS=1 BATCHES=262144 CHUNKS=64,128,32 ORDERS=1,0 SIGS=2,1 TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 tools/mgpu_probe.py > $O/mgpu_cfg4_n2.log 2> $O/mgpu_cfg4_n2.err
tail -3 $O/mgpu_cfg4_n2.err; cut -c1-400 $O/mgpu_cfg4_n2.log
