#!/usr/bin/env bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/last_smoke.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/last_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/last_pytest.log
tail -2 gpurun_out/last_smoke.log; tail -3 gpurun_out/last_pytest.log
