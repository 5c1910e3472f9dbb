run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 5 --no-train --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), [round(k['ms'],4) for k in d['kernels']])"; }
CTVQ_PEER_OVERLAP=0 run no_overlap
run overlap_512
CTVQ_PEER_THREADS=128 CTVQ_PEER_BLOCKS=1 run overlap_128x1
CTVQ_PEER_THREADS=64 CTVQ_PEER_BLOCKS=4 run overlap_64x4
