set -e
cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from __graft_entry__ import load_pkg
pkg=load_pkg()
g=np.load('tests/golden/frames_siso.npz')
rng=np.random.default_rng(21)
x=np.tile(g['iq'],2)
x=(x+(0.1875/np.sqrt(2*10**2.8))*(rng.standard_normal(x.size)+1j*rng.standard_normal(x.size))).astype(np.complex64)
rx=pkg.Receiver(device=0); preac,preconj=rx.presiso(x); rx.close()
os.makedirs('/tmp/rcd',exist_ok=True)
preac.tofile('/tmp/rcd/preac.f32'); preconj.astype(np.complex64).tofile('/tmp/rcd/preconj.c64'); x.tofile('/tmp/rcd/sig.c64')
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/blocks_launches.csv tests/gr_mock/build/run_chain 1 0 0 1 8192 0 /tmp/rcd /tmp/rcd/out.txt > gpurun_out/blocks_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/blocks_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0].replace('<unnamed>::','').replace('void ','')
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=float(r[-1].replace(',',''))
for k,v in agg.items(): print("%-40s %5d launches  %9.1f us total  %7.2f us avg"%(k,v[0],v[1]/1e3,v[1]/v[0]/1e3))
PY
