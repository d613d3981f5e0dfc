"""Stall breakdown of one profiled launch: `ncu -i X.ncu-rep --page raw --csv` + `--page source --csv` digests."""
import csv
import sys
from collections import Counter


def raw(path, idx=0):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2 + idx]
    for h, u, v in zip(hdr, units, data):
        if 'issue_stalled' in h and 'per_issue_active.ratio' in h:
            try:
                if float(v) > 0.1:
                    print(f"  {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {float(v):6.2f}")
            except ValueError:
                pass


def source(path, top=25, seg=0):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    rows = rows[starts[seg]:starts[seg + 1]]
    hdr = rows[1]; data = rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    ops = Counter(); opn = Counter(); opex = Counter(); st = {}
    for r in data:
        if len(r) <= ix['stall_wait']:
            continue
        src = r[ix['Source']].split()
        op = src[0] if not src[0].startswith('@') else src[1]
        parts = op.split('.')
        op = parts[0] + ('.' + '.'.join(parts[1:3]) if parts[0].startswith(('LD', 'ST', 'RED', 'ATOM')) else '')
        s = int(r[ix['# Samples']]); ops[op] += s; opn[op] += 1; opex[op] += int(r[ix['Instructions Executed']])
        d = st.setdefault(op, Counter())
        for k in stalls:
            d[k] += int(r[ix[k]])
    tot = sum(ops.values()); totex = sum(opex.values())
    print("samples", tot, "executed warp-instr", totex)
    for op, s in ops.most_common(top):
        d = st[op]; t3 = ', '.join(f"{k[6:]}={v}" for k, v in d.most_common(3))
        print(f"{op:18s} n={opn[op]:4d} exec={opex[op]/totex*100:5.1f}% samples={s/tot*100:5.1f}%  {t3}")


def lines(path, seg=0, top=22):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    rows = rows[starts[seg]:starts[seg + 1]]
    hdr = rows[1]; data = rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[ix['# Samples']]) for r in data if len(r) > ix['stall_wait'])
    best = sorted([r for r in data if len(r) > ix['stall_wait']], key=lambda r: -int(r[ix['# Samples']]))[:top]
    for r in best:
        d = {k: int(r[ix[k]]) for k in st}
        t2 = ', '.join(f"{k[6:]}={v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:2])
        print(f"{int(r[ix['# Samples']])/tot*100:5.1f}% ex={r[ix['Instructions Executed']]:>9s} {r[ix['Source']][:58]:58s} {t2}")


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        source(sys.argv[2], 25, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
        lines(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
