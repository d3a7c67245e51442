# ncu captures of everything that is not the class-0 fp32 solve kernel; usage: bash tools/profile_rest.sh <tag>
set -x
TAG=${1:-dev}
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
# class-0 fp64 (BASELINE configs[2]), class-2 fp32 (configs[3])
C2="python bench.py --config 2 --steps 2 --warmup 1 --sets 2 --lean --no-cpu-baseline"
$C2 > gpurun_out/plain_cfg2_$TAG.log 2>&1 && $NCU -k "regex:solve_kernel<double, .int.64," -s 1 -c 1 -o gpurun_out/prof_cfg2_f64_$TAG $C2 > gpurun_out/ncu_cfg2_$TAG.log 2>&1
C3="python bench.py --config 3 --steps 2 --warmup 1 --sets 2 --lean --no-cpu-baseline"
$C3 > gpurun_out/plain_cfg3_$TAG.log 2>&1 && $NCU -k "regex:solve_kernel<float, .int.192," -s 1 -c 1 -o gpurun_out/prof_cfg3_h30_$TAG $C3 > gpurun_out/ncu_cfg3_$TAG.log 2>&1
# standing gait: class 1 (n = 120)
C1="python bench.py --gait stand --steps 2 --warmup 1 --sets 2 --lean --no-cpu-baseline"
$C1 > gpurun_out/plain_stand_$TAG.log 2>&1 && $NCU -k "regex:solve_kernel<float, .int.128," -s 1 -c 1 -o gpurun_out/prof_stand_$TAG $C1 > gpurun_out/ncu_stand_$TAG.log 2>&1
# input kernels as HBM streams at 2^20 robots
S="python tools/assemble_gait_stream.py"
$S > gpurun_out/assemble_gait_stream_$TAG.json 2> gpurun_out/assemble_gait_stream_$TAG.err && $NCU -k "regex:mpcq_assemble_kernel|mpcq_gait_kernel" -s 6 -c 2 -o gpurun_out/prof_inputs_$TAG $S > gpurun_out/ncu_inputs_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep
