#!/bin/bash
# A/B of the grid barrier of the fused P-side kernel (BB_OPT_PSIDE_BARRIER: 1 release reduction + immediate poll, 0 fence + atomicAdd + fence)
mkdir -p gpurun_out
export BENCH_VALUED=0
timeout 900 python -m pytest tests/test_gpu_cg.py tests/test_gpu_multi.py tests/test_gpu_gibbs.py -m gpu -x -q > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log; tail -3 gpurun_out/b_pytest.log
run() {
    local name=$1 wl=$2 st=$3 wu=$4; shift 4
    env "$@" timeout 300 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline > gpurun_out/b_$name.log 2>&1
    echo "$name rc=$? $(python - gpurun_out/b_$name.log <<'P'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print('it/s %.2f e2e %.2f ms/step %.3f ncg %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['mean_n_cg_iter']))
P
)"
}
for rep in 1 2; do for b in 0 1; do run shard8_bar${b}_$rep C4shard8 40 5 BB_OPT_PSIDE_BARRIER=$b; done; done
for b in 0 1; do run c4_bar$b C4 20 5 BB_OPT_PSIDE_BARRIER=$b; done
for b in 0 1; do run c3_bar$b C3 50 10 BB_OPT_PSIDE_BARRIER=$b; done
