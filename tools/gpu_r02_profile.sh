#!/bin/bash
# tools/gpu_r02_profile.sh -- single-GPU profiles of round 2 (run under gpurun): launch list of a bench run, ncu --set full of
# the SpMV, the dot / vector kernels and the setup kernels; raw reports land in gpurun_out/, summaries are made by
# tools/ncu_summaries.py on the CPU side and committed under profiles/.
cd "$(dirname "$0")/.."
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file $O/r02_launches_bench_256.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r02_launches_bench_256.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 3 -c 1 -o $O/r02_prof_matvec python tools/run_cg.py --nx 256 --iters 6 > $O/r02_prof_matvec.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_dot_maps|k_dot_walk|k_xr_update|k_waxpby" -s 12 -c 8 -o $O/r02_prof_vec python tools/run_cg.py --nx 256 --iters 8 > $O/r02_prof_vec.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_structure|k_elements_tables|k_gather_rows|k_dirichlet|k_jacobian_tables" -c 8 -o $O/r02_prof_setup python tools/asm_probe.py --nx 256 --reps 1 > $O/r02_prof_setup.log 2>&1
for w in 24 32; do python tools/dot_stats.py --nx 256 --iters 50 --warps $w 2>&1 | head -1 | cut -c1-340; done > $O/r02_dot_warps.txt
ls -la $O/*.ncu-rep
