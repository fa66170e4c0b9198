#!/bin/bash
# Runs every measurement the files under profiles/ are made from, on one B200 (under gpurun):
#     gpurun --timeout 1800 -- 'bash tools/refresh_profiles.sh r02'
# then, back in the container:  python tools/summarize_profiles.py r02  and copy gpurun_out/bench_<tag>_*.json.
tag=${1:-r02}
only=${2:-all}     # `quick`: the legs whose code changed last (configs 2-4, records) and the captures
out=gpurun_out
mkdir -p $out
b() { name=$1; shift; timeout 600 python bench.py "$@" > $out/bench_${tag}_$name.json 2> $out/bench_${tag}_$name.err; echo "$name: $(cut -c1-160 $out/bench_${tag}_$name.json)"; }
b cfg2_n1 --steps 20 --warmup 3
[ $only = all ] && b cfg2_n1_reference_arm --impl reference --steps 3 --warmup 1
[ $only = all ] && b cfg2_n1_vector_stores --steps 10 --warmup 3 --no-e2e --no-cpu --no-extras --store-path stg
b cfg3_mixed_n1 --workload mixed_cfg3 --steps 10 --warmup 3 --no-e2e --no-extras
b cfg4_feasibility_n1 --workload montecarlo_cfg4 --n-per-gpu 10000000 --steps 5 --warmup 2 --no-e2e
[ $only = all ] && b f2_letters_T_n1 --workload letters_T --steps 10 --warmup 3 --no-e2e
[ $only = all ] && b f2_polyline_mix_n1 --workload polyline_mix --steps 10 --warmup 3 --no-e2e --no-cpu
b f3_records_n1 --records --steps 10 --warmup 3 --no-e2e
[ $only = all ] && b f3_records_letters_T_n1 --records --workload letters_T --steps 10 --warmup 3 --no-e2e
[ $only = all ] && b f4_transitions_n1 --workload transitions --steps 10 --warmup 3
[ $only = all ] && b f4_transitions_512k_n1 --workload transitions --n-per-gpu 524288 --steps 5 --warmup 2
# launch list of the default step, then one full capture of the dominant kernels (never a bench value)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --no-e2e --no-cpu --no-extras --steps 3 --warmup 2 > $out/ncu_ll.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eval_kernel|plan_phase_kernel" -c 6 \
    -o $out/${tag}_prof -f python bench.py --no-e2e --no-cpu --no-extras --steps 2 --warmup 2 --n-per-gpu 65536 > $out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 3 -c 1 -o $out/prof_rec_tma -f \
    python bench.py --records --steps 1 --warmup 1 --no-cpu --no-e2e > $out/ncu_rec.log 2>&1
[ $only = all ] && timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tgx::reduce_kernel -s 3 -c 1 -o $out/prof_feas -f \
    python bench.py --workload montecarlo_cfg4 --steps 2 --warmup 3 --no-cpu --no-e2e --n-per-gpu 2000000 > $out/ncu_feas.log 2>&1
[ $only = all ] && timeout 400 ncu --set full --clock-control none --import-source on -k regex:plan_fill_kernel -s 3 -c 1 -o $out/prof_plan_fill -f \
    python bench.py --workload montecarlo_cfg4 --steps 2 --warmup 3 --no-cpu --no-e2e --n-per-gpu 1000000 > $out/ncu_pf.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"tgx::eval_kernel|tgx::plan_phase_kernel" -s 2 -c 2 -o $out/prof_cfg3 -f \
    python bench.py --workload mixed_cfg3 --steps 1 --warmup 2 --no-cpu --no-e2e --no-extras --n-per-gpu 262144 > $out/ncu_cfg3.log 2>&1
[ $only = all ] && b cfg2_n1_pipeline --pipeline --steps 10 --warmup 3 --no-e2e --no-cpu --no-extras
ls -la $out | tail -8
