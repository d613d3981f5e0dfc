#!/bin/bash
# DRAM bytes + time of the fused-step launches of tools/ncu_target.py under the current environment
# usage: tools/dram_bytes.sh <tag>
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:"t3_kernel|p2_kernel|t4_kernel" -c 3 --csv python tools/ncu_target.py 1024 64 64 1 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
out={}
for r in rows[1:]:
    out.setdefault(r[ix['ID']],{})[r[ix['Metric Name']]]=(r[ix['Metric Value']],r[ix['Metric Unit']])
for k,v in out.items():
    print('$1', 'launch',k, ' '.join(f'{m.split(\"__\")[1][:18]}={a}{u}' for m,(a,u) in v.items()))
"
