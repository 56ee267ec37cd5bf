mkdir -p gpurun_out
python -m pytest tests/test_wdsp_gpu.py tests/test_wdsp_variants_gpu.py tests/test_wdsp_emnr_gpu.py tests/test_wdsp_snba_gpu.py tests/test_quisk_swapin_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t2.txt
for w in rxa_fm rxa_usb; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/b_$w.json 2> gpurun_out/b_$w.err; done
python bench.py --workload rxa_usb --channels 1024 --no-cpu-baseline > gpurun_out/b_rxa_usb_1024.json 2> gpurun_out/b_rxa_usb_1024.err
QUISK_FIR_MAC_SINGLE=1 python bench.py --workload rxa_usb --channels 1024 --no-cpu-baseline > gpurun_out/b_rxa_usb_1024_old.json 2>/dev/null
QUISK_FIR_MAC_SINGLE=1 python bench.py --workload rxa_fm --no-cpu-baseline > gpurun_out/b_rxa_fm_old.json 2>/dev/null
