# 2 GPUs: the multi-GPU parity program on plain row blocks and on compressed (row-pattern) row blocks, then the 2-GPU bench line
mkdir -p gpurun_out
T=${1:-r2mg}
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout -s KILL 400 $L --master-port 29511 tests/multi_gpu_check.py) > gpurun_out/${T}_mgc2.log 2>&1; echo "mgc rc=$?" >> gpurun_out/${T}_mgc2.log
(LCGB200_CHECK_COMPRESS=1 timeout -s KILL 400 $L --master-port 29512 tests/multi_gpu_check.py) > gpurun_out/${T}_mgc2_compressed.log 2>&1; echo "mgc rc=$?" >> gpurun_out/${T}_mgc2_compressed.log
(timeout -s KILL 400 $L --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/${T}_bench2.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench2.log
(timeout -s KILL 300 $L --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --compress --no-extra-legs) > gpurun_out/${T}_bench2_compressed.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench2_compressed.log
grep -c " OK" gpurun_out/${T}_mgc2.log; tail -2 gpurun_out/${T}_mgc2.log; grep -c " OK" gpurun_out/${T}_mgc2_compressed.log; tail -2 gpurun_out/${T}_mgc2_compressed.log
grep -o '"value": [0-9.]*' gpurun_out/${T}_bench2.log | head -1; grep -o '"value": [0-9.]*' gpurun_out/${T}_bench2_compressed.log | head -1
