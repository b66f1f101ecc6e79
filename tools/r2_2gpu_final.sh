#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp_gpu.py::test_cli_two_ranks -m gpu -q -rf -k "P2 or two_ranks" 2>&1 | tail -6 > gpurun_out/${1}_pytest_2gpu.log
cat gpurun_out/${1}_pytest_2gpu.log
