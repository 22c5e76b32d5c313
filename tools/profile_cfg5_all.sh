#!/bin/bash
# Evidence for the time-batched route on config 5: timelines of one call (needs the -DOHS_TB_TRACE
# library: OHS_NVCC_EXTRA=-DOHS_TB_TRACE ... build_library(out=open-headstage_b200/libohs_cuda_tbtrace.so)), the sweep of
# the EQ pre-pass's shape, blocks per call, the ncu launch list of a serialised call and full captures of its kernels.
T=$PWD/open-headstage_b200/libohs_cuda_tbtrace.so
OHS_LIB_OVERRIDE=$T python tools/trace_time_batched.py 256 2> gpurun_out/r02_timeline_config5_k256.txt >/dev/null
OHS_LIB_OVERRIDE=$T python tools/trace_time_batched.py 64 2> gpurun_out/r02_timeline_config5_k64.txt >/dev/null
OHS_TB_OVERLAP=0 OHS_LIB_OVERRIDE=$T python tools/trace_time_batched.py 64 2> gpurun_out/r02_timeline_config5_k64_serial.txt >/dev/null
tools/sweep_cfg5_prepass.sh > gpurun_out/r02_sweep_cfg5_prepass.txt 2>&1
for k in 64 128 256 512; do echo "K=$k overlapped: $(python tools/profile_cfg5_batched.py $k 2>&1 | tail -1)  serial: $(OHS_TB_OVERLAP=0 python tools/profile_cfg5_batched.py $k 2>&1 | tail -1)"; done > gpurun_out/r02_sweep_cfg5_blocks_per_call.txt
python tools/profile_cfg5_batched.py 64 > /dev/null 2>&1 && OHS_TB_OVERLAP=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_ncu_launches_config5_time_batched_k64_final.csv python tools/profile_cfg5_batched.py 64 > /dev/null 2>&1
for k in forward_kernel inverse_kernel bin_conv render_kernel; do OHS_TB_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r02_tb_final_$k -f python tools/profile_cfg5_batched.py 64 > /dev/null 2>&1; done
