#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -6 > gpurun_out/${1}_pytest.log
cat gpurun_out/${1}_pytest.log
