import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]; data=rows[2:]
want=['Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_shared_mem','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','lts__t_sector_hit_rate.pct','l1tex__t_bytes_pipe_lsu_mem_global_op_ldgsts_cache_access.sum']
idx={h:i for i,h in enumerate(hdr)}
for w in want:
    if w in idx:
        print("%-72s %-12s %s"%(w,units[idx[w]]," | ".join(r[idx[w]][:24] for r in data)))
