mkdir -p gpurun_out
(timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-ref-cuda) > gpurun_out/r2w_bench8.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_bench8.log
tail -c 600 gpurun_out/r2w_bench8.log
