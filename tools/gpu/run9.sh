timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q10_pytest.log 2>&1; tail -15 gpurun_out/q10_pytest.log
