# gpurun recipe behind profiles/: bench line, ncu launch list, one `ncu --set full` capture of the class-0 solve kernel
# usage (on the GPU box): bash tools/profile.sh <tag> [extra bench args]
set -x
TAG=${1:-dev}; shift
CMD="python bench.py --steps 2 --warmup 1 --sets 4 --no-cpu-baseline $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -20 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:solve_kernel<float, .int.64," -s 1 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
tail -3 gpurun_out/ncu_f_$TAG.log | cut -c1-300
