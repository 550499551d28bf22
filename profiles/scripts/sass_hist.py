import csv, sys, collections
f = sys.argv[1]
rows = list(csv.reader(open(f)))
hdr = rows[1]
iA, iS, iN, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
data = []
for r in rows[2:]:
    if len(r) <= iE: continue
    src = r[iS].strip(); 
    if src.startswith("@"): src = src.split(None, 1)[1]
    op = src.split()[0].split(".")[0] if src else "?"
    full = src.split()[0]
    e = int(r[iE] or 0); s = int(r[iN] or 0)
    ops[full if op in ("LDS","STS","LDG","STG","SHFL","MUFU","BAR","LD","ST") else op] += e; samp[op] += s; tot += e; tots += s
    data.append((e, s, src))
print("total executed warp-instr", tot, "samples", tots)
for k, v in ops.most_common(28): print(f"{k:14s} {v:14d} {100*v/tot:6.2f}%   samples {100*samp[k.split('.')[0]]/max(tots,1):6.2f}%")
if len(sys.argv) > 2:
    # cumulative profile by position: print blocks of N instructions with executed + samples
    N = int(sys.argv[2])
    for i in range(0, len(data), N):
        blk = data[i:i+N]
        e = sum(b[0] for b in blk); s = sum(b[1] for b in blk)
        print(f"{i:5d} exec {100*e/tot:5.1f}% samp {100*s/max(tots,1):5.1f}%  {blk[0][2][:50]}")
