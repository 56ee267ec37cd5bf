#!/bin/bash
# Round-end GPU pass: tests, smoke, benches and the ncu launch lists that go into profiles/.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_r1_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_ref.json 2>/dev/null
python bench.py --workload channelizer > gpurun_out/bench_r1_channelizer.json 2>/dev/null; tail -c 400 gpurun_out/bench_r1_channelizer.json
python bench.py --workload panadapter --channels 16 --block 1048576 --no-cpu-baseline > gpurun_out/bench_r1_pan16.json 2>/dev/null; tail -c 300 gpurun_out/bench_r1_pan16.json
python bench.py --workload rxa_usb --no-cpu-baseline > gpurun_out/bench_r1_rxa_usb.json 2>/dev/null
python bench.py --workload rxa_fm --no-cpu-baseline > gpurun_out/bench_r1_rxa_fm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_decim|rx_tail|nco_|polyfir|hb45|unpack|demod|tune_" -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pfb -c 8 --csv --log-file gpurun_out/launches_r1_channelizer.csv python bench.py --workload channelizer --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
tail -3 gpurun_out/launches_r1_channelizer.csv | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pan -c 12 --csv --log-file gpurun_out/launches_r1_panadapter.csv python bench.py --workload panadapter --channels 16 --block 1048576 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
python bench.py --nco closed --no-cpu-baseline > gpurun_out/bench_r1_n1_nco_closed.json 2>/dev/null
ncu --section SpeedOfLight --section WarpStateStats --section MemoryWorkloadAnalysis --section Occupancy --section SchedulerStats --section LaunchStats --clock-control none -k regex:pfb_f -c 2 -f -o gpurun_out/prof_pfb_r1 python tools/pfb_once.py 65536 0 > /dev/null 2>&1
ncu --section SpeedOfLight --section WarpStateStats --section MemoryWorkloadAnalysis --section Occupancy --section SchedulerStats --section LaunchStats --clock-control none -k regex:pan_accumulate -c 1 -f -o gpurun_out/prof_pan_r1 python bench.py --workload panadapter --channels 16 --block 1048576 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
ncu --set full --import-source on --clock-control none -k regex:fused_decim -s 2 -c 1 -f -o gpurun_out/prof_fused_r1 python bench.py --channels 1184 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
