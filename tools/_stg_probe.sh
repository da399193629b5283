cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
MILLIPYDE_GAUSS_COLUMN=fma timeout -s KILL 200 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__data_pipe_lsu_wavefronts.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:gauss_stream -s 3 -c 1 --csv --log-file gpurun_out/stg_fma.csv python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu --no-e2e > /dev/null 2>&1
cat gpurun_out/stg_fma.csv | tail -8
