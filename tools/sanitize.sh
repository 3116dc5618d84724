#!/bin/bash
# usage (GPU box, via gpurun): tools/sanitize.sh   -- compute-sanitizer over the small GPU tests (memcheck) and the smoke run (racecheck)
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_edge.py tests/test_gpu_slab.py -m gpu -q -x 2>&1 | tail -4
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(test_molecule_terms and fields) or (test_step_mc and lipo_eq) or (test_fused_step_kernel_with_every and bead2)" 2>&1 | tail -4
timeout 120 compute-sanitizer --tool racecheck --print-limit 3 python __graft_entry__.py smoke 2>&1 | tail -3
