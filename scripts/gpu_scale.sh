set -x
N=${1:-8}
nvidia-smi -L | head -8
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json; cat gpurun_out/scale_$N.json; tail -3 gpurun_out/scale_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 bench.py --impl reference --gpus $N --steps 2 --warmup 0 2>/dev/null | tail -1
