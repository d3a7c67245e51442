set -x
TAG=r02v8
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
for c in 0 2 3 4; do python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg$c.json 2> gpurun_out/bench_${TAG}_cfg$c.err; done
bash tools/profile.sh $TAG
python tools/assemble_gait_stream.py > gpurun_out/assemble_gait_stream_$TAG.json 2> gpurun_out/assemble_gait_stream_$TAG.err
C3="python bench.py --config 3 --steps 2 --warmup 1 --sets 2 --lean --no-cpu-baseline"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k "regex:solve_kernel<float, .int.192," -s 1 -c 1 -o gpurun_out/prof_cfg3_h30_$TAG $C3 > gpurun_out/ncu_cfg3_$TAG.log 2>&1
for f in "" _cfg0 _cfg2 _cfg3 _cfg4; do python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}$f.json')); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('e2e_pipelined',{}).get('value'), d['roofline']['frac'], d.get('latency'))"; done
