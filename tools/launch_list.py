"""Per-launch durations of an `ncu --metrics gpu__time_duration.sum --csv` log, in launch order:
    python tools/launch_list.py <csv>"""
import csv,sys
from collections import OrderedDict
rows=[r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
d=OrderedDict()
for r in rows:
    d.setdefault((int(r[0]), r[4][:60]),{})[r[12]]=r[14]
tot=0
for k,v in d.items():
    us=float(v['gpu__time_duration.sum'].replace(',',''))/1e3; tot+=us
    print(k[0], k[1], "us=%.1f"%us)
print("total us", tot)
