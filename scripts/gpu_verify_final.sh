#!/bin/bash
# last check of the round's final build: the whole GPU suite, smoke(), and the default bench line
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/v_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/v_bench_c4.log 2>&1
tail -n 4 gpurun_out/v_pytest.log; tail -n 2 gpurun_out/v_smoke.log; grep '^{' gpurun_out/v_bench_c4.log | tail -1 | cut -c1-300
