import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); h,u=rows[0],rows[1]
for d in rows[2:]:
    print("##", d[h.index("Kernel Name")][:80])
    for k in ('l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed','l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum','lts__t_sectors_srcunit_tex_op_read.sum','smsp__thread_inst_executed_per_inst_executed.ratio','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.avg','smsp__cycles_active.avg'):
        if k in h: print(f"{k:85s} {d[h.index(k)]:>20s} {u[h.index(k)]}")
