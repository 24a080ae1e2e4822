mkdir -p gpurun_out
(timeout 600 python tools/exact_probe.py) > gpurun_out/r2p_exact.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_exact.log
(timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r2p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_pytest.log
tail -5 gpurun_out/r2p_pytest.log; grep -c "^OK" gpurun_out/r2p_exact.log; tail -3 gpurun_out/r2p_exact.log
