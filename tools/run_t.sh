mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py) > gpurun_out/r2t_mgc2.log 2>&1; echo "mgc rc=$?" >> gpurun_out/r2t_mgc2.log
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu --no-ref-cuda) > gpurun_out/r2t_bench2.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_bench2.log
tail -4 gpurun_out/r2t_mgc2.log; tail -c 300 gpurun_out/r2t_bench2.log
