cd /root/repo
timeout 900 python -m pytest tests/ -q -m gpu > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/r02w_pytest.log
tail -8 gpurun_out/r02w_pytest.log
