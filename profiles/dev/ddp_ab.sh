# A/B of the gradient-exchange modes at 2 GPUs against the 1-GPU step (gpurun --gpus 2 -- bash profiles/dev/ddp_ab.sh)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', d['ms_per_step'], d.get('ddp_selfcheck',{}).get('max_rel_grad_err'))"; }
LSTHM_DDP_OVERLAP=0 LSTHM_DDP_BUCKET_MB=64 run 29621 "no-overlap-1bucket"
LSTHM_DDP_OVERLAP=1 LSTHM_DDP_BUCKET_MB=2 run 29624 "overlap-2MB"
LSTHM_DDP_OVERLAP=1 LSTHM_DDP_BUCKET_MB=1 run 29625 "overlap-1MB"
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['ms_per_step'])"
python -m pytest tests/test_ddp_nccl_gpu.py tests/test_adam_gpu.py -m gpu -x -q 2>&1 | tail -2
