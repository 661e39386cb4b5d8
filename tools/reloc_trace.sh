LOCREG_PROFILE_TRACE=1 HYP=2048 python tools/reloc_breakdown.py 2>&1 | grep "locreg-trace" > gpurun_out/r2s_reloc_trace.txt; wc -l gpurun_out/r2s_reloc_trace.txt
