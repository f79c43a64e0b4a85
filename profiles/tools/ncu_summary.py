import csv, subprocess
rows=list(csv.reader(open("/tmp/hot6_raw.csv")))
hdr=rows[0]; units=rows[1]
cols=[("gpu__time_duration.sum","time us"),("launch__registers_per_thread","regs"),("Grid Size","grid"),("Block Size","block"),
("launch__waves_per_multiprocessor","waves/SM"),("sm__warps_active.avg.pct_of_peak_sustained_active","warps active %"),
("smsp__inst_executed.sum","warp instructions"),("smsp__thread_inst_executed_per_inst_executed.ratio","threads/inst"),
("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","fp64 pipe %"),("dram__bytes_read.sum","dram read MB"),("dram__bytes_write.sum","dram write MB"),
("smsp__warps_eligible.avg.per_cycle_active","eligible warps/cycle"),("sm__throughput.avg.pct_of_peak_sustained_elapsed","sm throughput %")]
print("# ncu --set full, one launch of each hot kernel (sim50x128, real input, session-6 build: counting sort, one split-time body, convergent continued fraction; fast path, 4 pairs per warp, two chain groups)\n")
print("Command: `ncu --set full --clock-control none --import-source on -k regex:'k_move|k_weigh|k_accept|k_split_t_fast|k_swap|k_changeu' -s 2108 -c 7 -o gpurun_out/r2s6_hot python profiles/tools/one_step.py sim50x128 300 3 1 4` (profiles/tools/r2_gpu20.sh).  The engine runs two chain groups: every launch covers 64 chains = 3,200 pairs; per step there are two of each.  Times under ncu are cold-cache and serialised; the launch list of the same program without the full sections is `r2s6_launches_sim50x128.csv`.\n")
print("| kernel | "+" | ".join(c[1] for c in cols)+" |")
print("|---|"+"---|"*len(cols))
ki=hdr.index("Kernel Name")
seen=set()
for r in rows[2:]:
    k=r[ki].split("(")[0].replace("void ","").strip()
    if k in seen: continue
    seen.add(k)
    vals=[]
    for c,_ in cols:
        v=r[hdr.index(c)]
        try: v="%.4g"%float(v.replace(",",""))
        except: pass
        vals.append(v)
    print("| `%s` | "%k+" | ".join(vals)+" |")
print()
cand=[h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print("Top warp-stall reasons (warps stalled per issue-active cycle):\n")
seen=set()
for r in rows[2:]:
    k=r[ki].split("(")[0].replace("void ","").strip()
    if k in seen: continue
    seen.add(k)
    vals=[]
    for h in cand:
        try: vals.append((float(r[hdr.index(h)]), h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio","")))
        except: pass
    vals.sort(reverse=True)
    print("* `%s`: "%k+", ".join("%s %.2f"%(n,v) for v,n in vals[:6]))
print()
for kern,label in [("k_split_t_fast","k_split_t_fast"),("k_weigh","k_weigh"),("k_move","k_move"),("k_accept<","k_accept"),("k_accept_t","k_accept_t")]:
    out=subprocess.run(["python","profiles/tools/ncu_funcs.py","gpurun_out/r2s6_hot.ncu-rep",kern,"14"],capture_output=True,text=True).stdout
    print("## `%s` by enclosing source function (warp instructions executed, stall samples)\n\n```\n%s```\n"%(label,out))
