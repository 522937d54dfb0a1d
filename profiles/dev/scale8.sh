# 8-GPU run of the bench (both gradient-exchange modes) + 1-GPU run on the same box: gpurun --gpus 8 -- bash profiles/dev/scale8.sh
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 3 --no-extras --no-cpu-baseline $3 2>gpurun_out/s3_scale_n${N}_$2.err | tee gpurun_out/s3_scale_n${N}_$2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', d['n_gpus'], d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), d.get('ddp_selfcheck',{}).get('max_rel_grad_err'))"; }
LSTHM_DDP_OVERLAP=1 LSTHM_DDP_BUCKET_MB=2 run 29631 overlap2mb
LSTHM_DDP_OVERLAP=0 LSTHM_DDP_BUCKET_MB=64 run 29632 nooverlap1 --no-e2e
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tee gpurun_out/s3_scale_n1_ref.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['ms_per_step'], d['value'], d['e2e']['value'])"
