import json, sys
d = json.load(open(sys.argv[1]))
print('step ms (events)', round(d['step_ms_events'], 2), 'ms/step', round(d['ms_per_step'], 2))
for r in d['by_name'][:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"{r['name']:28s} calls {r['calls']:4d} ms {r['ms']:8.2f} GB/s {r['bytes']/max(r['ms'],1e-9)/1e6:8.1f} TF/s {r['flops']/max(r['ms'],1e-9)/1e9:7.1f}")
print()
for r in d['rows'][:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    print(f"{r['name']:28s} {r['key']:32s} calls {r['calls']:3d} ms {r['ms']:7.2f} GB/s {r['bytes']/max(r['ms'],1e-9)/1e6:8.1f} TF/s {r['flops']/max(r['ms'],1e-9)/1e9:7.1f}")
