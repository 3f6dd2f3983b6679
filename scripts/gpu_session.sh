#!/bin/bash
# full GPU suite + smoke with the final build
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/s16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s16_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/s16_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s16_smoke.log
tail -6 gpurun_out/s16_pytest.log; tail -3 gpurun_out/s16_smoke.log
