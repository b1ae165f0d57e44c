#!/bin/bash
# Profiling evidence of round 2 (run on the GPU box through gpurun; outputs land in gpurun_out/, the summaries are copied
# to profiles/ by hand).  Numbers printed under ncu / compute-sanitizer are never bench values.
set -x
O=gpurun_out
# 1. launch list of the bench command (kernel shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench512.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs > $O/r2_launches_bench512.out 2>&1
# 2. full capture of the dominant kernel (one-pass column kernel, whole 512^3 grid) and of the split-operand indexed kernel
ncu --set full --clock-control none --import-source on -k regex:query_col_kernel -c 1 -f -o $O/r2_col512 python scripts/grid_once.py 512 fp16 > $O/r2_col512.out 2>&1
ncu --set full --clock-control none -k regex:query_col_kernel --launch-skip 1 -c 1 -f -o $O/r2_col512_refine python scripts/grid_once.py 512 fp16r > $O/r2_col512_refine.out 2>&1
# 3. full capture of the marching-cubes kernels (HR volume: 8 launches)
ncu --set full --clock-control none -k regex:mc_ -c 8 -f -o $O/r2_mc512 python scripts/mc_profile.py 512 1 > $O/r2_mc512.out 2>&1
# 4. sanitizers on the smoke invocation (query, dense default precision, octree, marching cubes)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/r2_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/r2_sanitizer_racecheck.txt
tail -5 $O/r2_sanitizer_memcheck.txt $O/r2_sanitizer_racecheck.txt
ls -la $O/*.ncu-rep
