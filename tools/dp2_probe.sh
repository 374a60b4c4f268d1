run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
echo chunk3; run
echo chunk1; MTUS_DP_BLOCKS_PER_CHUNK=1 run
echo chunk6; MTUS_DP_BLOCKS_PER_CHUNK=6 run
echo chunk100; MTUS_DP_BLOCKS_PER_CHUNK=100 run
