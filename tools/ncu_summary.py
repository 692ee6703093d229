import sys,subprocess,csv,re
rep,title=sys.argv[1],sys.argv[2]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); h,u=rows[0],rows[1]
pats=[r'^gpu__time_duration.sum$',r'^dram__bytes_(read|write).sum$',r'gpu__dram_throughput.avg.pct',r'^sm__throughput.avg.pct',r'smsp__issue_active.avg.pct',r'sm__warps_active.avg.pct',r'sm__inst_executed_pipe_(alu|fma|lsu|fp64).avg.pct_of_peak_sustained_active',r'lsu_wavefronts_mem_shared.sum.pct',r'bank_conflicts_pipe_lsu_mem_shared.sum$',r'^smsp__inst_executed.sum$',r'launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit)',r'issue_stalled_.*per_issue_active']
print("## %s\n"%title)
for r in rows[2:]:
    print("`%s`\n"%r[h.index("Kernel Name")][:120])
    print("| counter | value | unit |\n|---|---|---|")
    for k,un,x in zip(h,u,r):
        if x and any(re.search(p,k) for p in pats) and 'not_issued' not in k and 'allocated' not in k:
            print("| %s | %s | %s |"%(k,x,un))
    print()
