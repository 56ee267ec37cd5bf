mkdir -p gpurun_out
python -m pytest tests/test_wdsp_gpu.py tests/test_wdsp_variants_gpu.py tests/test_quisk_swapin_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t5.txt
for f in stream group; do
QUISK_FIR_MAC=$f python bench.py --workload rxa_usb --channels 1024 --no-cpu-baseline > gpurun_out/b5_usb1024_$f.json 2> gpurun_out/b5.err
QUISK_FIR_MAC=$f python bench.py --workload rxa_fm --no-cpu-baseline > gpurun_out/b5_fm_$f.json 2>> gpurun_out/b5.err
QUISK_FIR_MAC=$f python bench.py --workload rxa_usb --no-cpu-baseline > gpurun_out/b5_usb_$f.json 2>> gpurun_out/b5.err
done
