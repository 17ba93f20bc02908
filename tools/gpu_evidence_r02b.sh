#!/bin/bash
# round 2, final evidence call: sanitizer on smoke() (memcheck / racecheck / synccheck; covers the chained forward, the
# tensor-core head kernels, the batched reductions), ncu --set full of the new kernels, batch-1 launch list
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.txt 2>&1; tail -3 gpurun_out/sanitizer_memcheck_smoke.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke.txt 2>&1; tail -3 gpurun_out/sanitizer_racecheck_smoke.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_synccheck_smoke.txt 2>&1; tail -3 gpurun_out/sanitizer_synccheck_smoke.txt
bash tools/gpu_prof_kernel.sh head_bwd_tc_kernel 1 prof_r02_head_bwd_tc
bash tools/gpu_prof_kernel.sh head_tapdot_kernel 0 prof_r02_head_tapdot
bash tools/gpu_prof_kernel.sh "conv64_tc_kernel" 22 prof_r02b_conv64_fwd_hr
