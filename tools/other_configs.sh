python bench.py --robot AliengoConfig --gait mix --dtype f64 --envs 16384 --sets 16 --steps 50 --no-cpu-baseline > gpurun_out/bench_v15_cfg3.json 2> gpurun_out/bench_v15_cfg3.err
python bench.py --horizon 30 --envs 4096 --sets 16 --steps 20 --no-cpu-baseline > gpurun_out/bench_v15_cfg4.json 2> gpurun_out/bench_v15_cfg4.err
python bench.py --envs 262144 --sets 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v15_cfg5_1gpu.json 2> gpurun_out/bench_v15_cfg5_1gpu.err
python bench.py --gait stand --envs 4096 --sets 16 --steps 50 --no-cpu-baseline > gpurun_out/bench_v15_stand.json 2> gpurun_out/bench_v15_stand.err
for f in cfg3 cfg4 cfg5_1gpu stand; do python -c "
import json,sys; d=json.load(open('gpurun_out/bench_v15_$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['solver'], d['config']['l2'])" || tail -3 gpurun_out/bench_v15_$f.err; done
