set -x
O=gpurun_out/r2k
mkdir -p $O
for MC in 0 1; do
MC=$MC SCALE=24 MODEL=5 DIM=128 BS=1 BATCHES=262144 CHUNKS=0 ORDERS=1 SIGS=2 FREE=0 TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29713 tools/mgpu_probe.py > $O/mgpu_cfg4_n4_mc$MC.log 2> $O/mgpu_cfg4_n4_mc$MC.err; tail -2 $O/mgpu_cfg4_n4_mc$MC.err; grep -v trace_rank $O/mgpu_cfg4_n4_mc$MC.log | cut -c1-330; grep '"trace_rank": 0' $O/mgpu_cfg4_n4_mc$MC.log | cut -c1-200
done
