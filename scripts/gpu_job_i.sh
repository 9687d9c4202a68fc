#!/usr/bin/env bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/i_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/i_pytest.log
tail -4 gpurun_out/i_smoke.log; tail -6 gpurun_out/i_pytest.log
