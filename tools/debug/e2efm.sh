mkdir -p gpurun_out
python bench.py --workload rxa_fm --no-cpu-baseline > gpurun_out/e2e_new.json 2>/dev/null
QUISK_DECIM_RB_GENERIC=1 python bench.py --workload rxa_fm --no-cpu-baseline > gpurun_out/e2e_old.json 2>/dev/null
python bench.py --workload rxa_fm --no-cpu-baseline > gpurun_out/e2e_new2.json 2>/dev/null
