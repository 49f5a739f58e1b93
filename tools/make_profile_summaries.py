import csv, json, re
from collections import defaultdict
bench=json.load(open('profiles/r2_bench_n1.json'))
ph=bench['phase_ms_per_step']; one=bench['ms_per_step_one_in_flight']
rows=[r for r in csv.reader(l for l in open('profiles/r2_launches.csv') if l.startswith('"'))]
hdr=rows[0]; ki,vi=hdr.index('Kernel Name'),hdr.index('Metric Value')
fam=defaultdict(lambda:[0,0.0])
def family(n):
    n=re.sub(r'\(.*','',n).replace('void ','')
    n=re.sub(r'<.*','',n)
    if n.startswith('eval_check_rv32im'): return 'eval_check_rv32im (36 parts)'
    return n
for r in rows[1:]:
    f=family(r[ki]); fam[f][0]+=1; fam[f][1]+=float(r[vi].replace(',',''))/1e6
tot=sum(v[1] for v in fam.values()); n=sum(v[0] for v in fam.values())
phase_of={'eval_check_rv32im (36 parts)':'eval_check','r0::p2_hash_rows_kernel':'hash_rows','r0::p2_hash_fold_kernel':'hash_fold','r0wg::k_step_exec':'witgen','r0wg::k_step_accum':'accum','r0::bit_reverse_tma_kernel':'bit_reverse'}
L=[]
L.append('# Round 2 - ncu launch list of one whole prove_core at po2 = 20 (real loop-guest segment, witness generated on the device)\n')
L.append('Command (after the same command had exited 0 without ncu in the same gpurun call, `tools/final_profile.sh`):\n')
L.append('    ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/<tag>_launches.csv python tools/profile_target.py\n')
L.append('Raw list: `profiles/r2_launches.csv` (%d launches = exactly one step, no warm-up proof). ncu serialises launches and runs them cold-cache: compare SHARES with the bench line\'s one-in-flight `phase_ms_per_step` (`profiles/r2_bench_n1.json`, %.1f ms per step one segment at a time), not absolutes.\n' % (n, one))
L.append('| kernel family | launches | device ms | share | bench phase (ms, share of %.1f) |'%one)
L.append('|---|---|---|---|---|')
for f,(c,ms) in sorted(fam.items(), key=lambda kv:-kv[1][1]):
    p=phase_of.get(f); extra='%s %.1f, %.1f %%'%(p,ph[p],100*ph[p]/one) if p else ''
    if ms<0.02: continue
    L.append('| `%s` | %d | %.3f | %.1f %% | %s |'%(f,c,ms,100*ms/tot,extra))
L.append('\ntotal device time in listed launches: %.2f ms over %d launches (bench: %.1f ms per step one at a time, %.1f with three segments in flight)'%(tot,n,one,bench['ms_per_step']))
open('profiles/r2_launches_summary.md','w').write('\n'.join(L)+'\n')

# ---- sections
sec=json.load(open('profiles/r2_ncu_sections.json'))
def fmt(rec,keys):
    return ' | '.join(('%.3f'%rec[k] if k=='ms' else ('%.2f'%rec[k] if k in ('dram_GB','long_scoreboard','math_pipe_throttle','no_instruction','wait','barrier','not_selected') else '%.1f'%rec[k])) if isinstance(rec.get(k),(int,float)) else '-' for k in keys)
ec=[r for r in sec if r['kernel'].startswith('eval_check_rv32im_p')]
ec.sort(key=lambda r:int(r['kernel'].split('_p')[1]))
keys=['ms','dram_GB','dram_pct','fmaheavy','alu','issue','warps','regs','l1hit','l2hit','long_scoreboard','math_pipe_throttle','no_instruction','wait']
M=[]
M.append('# Round 2 - ncu capture of one whole prove_core at po2 = 20 (final kernels)\n')
M.append('Command (same gpurun call, after the plain run of the same command had exited 0; `tools/final_profile.sh`):\n')
M.append('    ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section Occupancy \\\n        --section LaunchStats --section WarpStateStats --section SchedulerStats --section InstructionStats --clock-control none \\\n        -k "regex:eval_check_rv32im_p|p2_hash_rows_kernel|p2_hash_fold_kernel|k_step_exec|k_step_accum|ntt_strided|ntt_fwd_contig|ntt_inv_contig|bit_reverse_tma" -c 90 python tools/profile_target.py\n')
M.append('The `.ncu-rep` holds 36 embedded cubins and is over the 64 MiB that travels back from the GPU box, so it was exported there with\n`ncu -i ... --page raw --csv`; `tools/ncu_sections_summary.py` turns that export into `profiles/r2_ncu_sections.json` (all %d captured launches) and the\ntables below. DRAM bytes = `dram__bytes.sum.per_second` x `gpu__time_duration.sum`. Stall columns are `smsp__average_warps_issue_stalled_*_per_issue_active`.\n'%len(sec))
M.append('## eval_check: all 36 generated part kernels after the per-part tuning (37 % of the step)\n')
M.append('| part | '+' | '.join(keys)+' |'); M.append('|---|'+'---|'*len(keys))
for r in ec: M.append('| p%s | '%r['kernel'].split('_p')[1]+fmt(r,keys)+' |')
tms=sum(r['ms'] for r in ec); tgb=sum(r['dram_GB'] for r in ec)
def w(k): return sum(r[k]*r['ms'] for r in ec)/tms
M.append('\nSum: %.2f ms, DRAM traffic **%.1f GB per proof = %.1fx the algorithmic 5.35 GB** (every part re-reads its tap columns; %.3f GB per launch - this is `roofline.traffic`, `profiles/r2_dram_traffic.json`).'%(tms,tgb,tgb/5.35,tgb/36))
M.append('Time-weighted: `sm__pipe_fmaheavy_cycles_active` **%.1f %%**, alu %.1f %%, issue slots %.1f %%, warps active %.1f %%, DRAM %.1f %%; stalls per issue: no_instruction %.2f, long_scoreboard %.2f, wait %.2f, math_pipe_throttle %.2f.'%(w('fmaheavy'),w('alu'),w('issue'),w('warps'),w('dram_pct'),w('no_instruction'),w('long_scoreboard'),w('wait'),w('math_pipe_throttle')))
M.append('Reading: both integer pipes are loaded alike (the alu counter does not include the 2 cycles IMAD.WIDE / IMAD.HI hold it, DESIGN.md 3); the largest stall is instruction fetch of the straight-line code, then tap latency. Before the tuning (`git show 9ccb03a:profiles/r2_ncu_summaries.md`): 51.8 ms, fmaheavy 68.3 %.\n')
M.append('## Other kernels of the step\n')
keys2=keys+['barrier','not_selected']
M.append('| kernel | '+' | '.join(keys2)+' |'); M.append('|---|'+'---|'*len(keys2))
for r in sec:
    if r['kernel'].startswith('eval_check_rv32im_p'): continue
    M.append('| `%s` | '%r['kernel']+fmt(r,keys2)+' |')
M.append('')
open('profiles/r2_ncu_summaries.md','w').write('\n'.join(M)+'\n')
hr=[r for r in sec if 'p2_hash_rows' in r['kernel']]
json.dump({'eval_check':{'dram_bytes_per_launch_po2_20':tgb*1e9/36,'launches':36,'dram_bytes_per_proof_po2_20':tgb*1e9,'source':'profiles/r2_ncu_sections.json (ncu sections capture of one prove_core at po2=20, dram__bytes.sum.per_second x gpu__time_duration)'},
 'hash_rows':{'dram_bytes_per_launch_po2_20':sum(r['dram_GB'] for r in hr)*1e9/max(1,len(hr)),'launches':len(hr)}},open('profiles/r2_dram_traffic.json','w'),indent=1)
print(tms,tgb,w('fmaheavy'),w('alu'))
