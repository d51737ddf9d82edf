"""Summarise an `ncu --page source --csv` export: stall samples per SASS region between marker
instructions, top lines.  usage: python scripts/ncu_roles.py src.csv [threshold]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
th = int(sys.argv[2]) if len(sys.argv) > 2 else 300
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
S = lambda i: int(data[i][ix['# Samples']])
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('total samples', sum(S(i) for i in range(len(data))), 'warp instr',
      sum(int(r[ix['Instructions Executed']]) for r in data))
for i, r in enumerate(data):
    src = r[ix['Source']].strip()
    mark = re.search(r'UTCHMMA|LDGSTS|UTMALDG|STTM|LDTM|TRYWAIT|ATOMG|UTCBAR|LDS\.128|STG|EXIT', src)
    if S(i) >= th or (mark and S(i) >= 30):
        st = {h: int(r[ix[h]]) for h in stall_cols if int(r[ix[h]]) > 0}
        st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(i, S(i), r[ix['Instructions Executed']], src[:64], st)
