import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        if line: print(line)
        continue
    d=json.loads(line)
    k=d.get('kernels',{})
    print(f"value={d['value']/1e9:.1f} Grows/s ms/step={d['ms_per_step']:.2f}", {n:(round(v['ms_total']/d['steps'],2), round(v['achieved_gbs'])) for n,v in k.items()})
