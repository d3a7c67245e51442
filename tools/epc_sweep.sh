# resident one-warp robots per CTA (class 0): rebuild with -DMPCQ_EPC0=k and time 4 096 / 65 536 robots
for k in "$@"; do
  (cd pympc_quadruped_b200/csrc && MPCQ_EXTRA_NVCC_FLAGS=-DMPCQ_EPC0=$k python build.py --force >/dev/null) && echo "EPC0=$k" && python tools/ctas_sweep.py 8 2>&1 | tail -2
done
