#!/bin/bash
# weak scaling on one 8-GPU box (gpurun --gpus 8 -- 'bash tools/gpu_scale8.sh'): N = 1, 4, 8 back to back, then the sharded inference sweep
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "n1 exit $?"
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "n$n exit $?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload infer_c5 > gpurun_out/scale_infer_n8.json 2> gpurun_out/scale_infer_n8.err; echo "infer n8 exit $?"
grep -h '^{' gpurun_out/scale_n*.json | cut -c1-200
