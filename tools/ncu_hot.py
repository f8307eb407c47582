"""Print headline metrics + hottest SASS lines (stall samples) of an .ncu-rep."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum ', 'sm__inst_executed.sum ', 'launch__grid_size',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct']
hdr = rows[0]
for vals in rows[2:]:
    if len(vals) != len(hdr):
        continue
    print('kernel:', vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?', ' grid', vals[hdr.index('Grid Size')] if 'Grid Size' in hdr else '?',
          ' block', vals[hdr.index('Block Size')] if 'Block Size' in hdr else '?')
    for i, h in enumerate(hdr):
        if any(h == k.strip() or (k.endswith('limit') and k in h) for k in keys) or h.startswith('sm__pipe_tensor') \
                or h.startswith('sm__pipe_tc') or h.startswith('sm__inst_executed_pipe_tensor') or h.startswith('dram__throughput') \
                or h.startswith('smsp__inst_executed_pipe_tc') or h.startswith('sm__pipe_xu_cycles_active'):
            print(f'  {h} = {vals[i]} {rows[1][i]}')
vals = rows[2]
st = []
for i, h in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
        try: st.append((float(vals[i].replace(',', '')), h.split('stalled_')[1]))
        except Exception: pass
st.sort(reverse=True)
print('stalls:', ', '.join(f'{n}={int(v)}' for v, n in st[:8]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
col, ex = ci['Warp Stall Sampling (All Samples)'], ci['Instructions Executed']
rs = []
for idx, r in enumerate(rows[2:]):
    try: rs.append((int(r[col]), idx, r[ci['Source']].strip()[:100], r[ex]))
    except Exception: pass
tot = sum(v for v, *_ in rs)
print('total samples', tot, 'instructions', len(rs))
for v, idx, s, e in sorted(rs, reverse=True)[:top]:
    print(f'{v:6d} {100*v/max(tot,1):5.1f}%  #{idx:4d} x{e:>8s}  {s}')
