set -x
python bench.py > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v11.json 2> gpurun_out/bench_ref_v11.err
CMD="python bench.py --steps 2 --warmup 1 --sets 4 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v11.csv $CMD > gpurun_out/ncu14.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mpcq_solve_kernel -s 2 -c 1 -o gpurun_out/prof_r1_v11 $CMD > gpurun_out/ncu15.log 2>&1
tail -3 gpurun_out/ncu15.log | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/bench_v11.json'))
print({k:d[k] for k in ('value','ms_per_step','latency_ms','clocks','cpu_baseline','roofline')})
print(json.load(open('gpurun_out/bench_ref_v11.json'))['value'])
"
