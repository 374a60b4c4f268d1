run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NP:-2} --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"; }
echo default; run
echo nch4; NCCL_MAX_NCHANNELS=4 run
echo nch8; NCCL_MAX_NCHANNELS=8 run
echo sms140; MTUS_GEMM_SMS=140 run
