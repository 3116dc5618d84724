#!/bin/bash
# usage (GPU box, via gpurun): tools/sanitize.sh [tag]  -- compute-sanitizer over the small GPU tests (memcheck) and the smoke run
# (racecheck); full logs -> gpurun_out/sanitize_<tag>_*.log (copied to profiles/ by hand), summary lines on stdout
tag=${1:-r02}
mkdir -p gpurun_out
m() { n=$1; shift; timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 "$@" > gpurun_out/sanitize_${tag}_$n.log 2>&1; echo "memcheck $n: exit $? -- $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_${tag}_$n.log | tr '\n' ' ')"; }
m edge_slab python -m pytest tests/test_gpu_edge.py tests/test_gpu_slab.py -m gpu -q -x
m parity python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(test_molecule_terms and fields) or (test_step_mc and lipo_eq) or (test_fused_step_kernel_with_every and bead2) or (test_split_pair_engine and (gas or lipo_eq))"
m observables_substrate python -m pytest tests/test_gpu_observables.py tests/test_gpu_substrate.py -m gpu -q -x -k "not md_b200"
timeout 200 compute-sanitizer --tool racecheck --print-limit 3 python __graft_entry__.py smoke > gpurun_out/sanitize_${tag}_racecheck_smoke.log 2>&1
echo "racecheck smoke: exit $? -- $(grep -E 'RACECHECK SUMMARY|smoke ok' gpurun_out/sanitize_${tag}_racecheck_smoke.log | tr '\n' ' ')"
