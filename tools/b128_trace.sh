mkdir -p gpurun_out
DX_PROF_DUMP=1 python tools/b128_gemm_dump.py 2> gpurun_out/b128_dump.err | tee gpurun_out/b128_dump.log
python tools/gemm_dump_summary.py gpurun_out/b128_dump.err | head -50 | tee -a gpurun_out/b128_dump.log
DX_TC_DEBUG=1 python tools/x3_probe.py one 128 1536 512 > gpurun_out/x3_trace_128.log 2>&1; tail -40 gpurun_out/x3_trace_128.log
