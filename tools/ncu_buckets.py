"""Development tool: aggregate an `ncu --page source --csv` dump by code position and opcode."""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=[]
for r in rows[2:]:
    if len(r)<len(hdr) or not r[ix['# Samples']].isdigit(): continue
    data.append((int(r[ix['Address']],16), r[ix['Source']].strip(), int(r[ix['# Samples']]), int(r[ix['Instructions Executed']]), int(r[ix['stall_no_inst']]), int(r[ix['stall_long_sb']])))
base=data[0][0]
B=int(sys.argv[2]) if len(sys.argv)>2 else 1024
print("static instrs", len(data))
print("bucket  instr_exec(M)  samples  no_inst  long_sb  addr")
for b in range(0,len(data),B):
    chunk=data[b:b+B]
    ex=sum(c[3] for c in chunk); s=sum(c[2] for c in chunk); ni=sum(c[4] for c in chunk); ls=sum(c[5] for c in chunk)
    print(f"{b:6d} {ex/1e6:10.1f} {s:9d} {ni:9d} {ls:8d}  +{chunk[0][0]-base:#x}")
h=collections.Counter()
for a,src,s,ex,ni,ls in data:
    t=src.split()
    op=t[1] if t[0].startswith('@') else t[0]
    h[op.split('.')[0]]+=ex
tot=sum(h.values())
print("opcode mix (executed warp instructions):")
for k,v in h.most_common(22): print("  ",k, '%.1f%%'%(100*v/tot))
