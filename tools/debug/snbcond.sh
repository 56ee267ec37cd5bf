F=oracle/_ref/quisk_full
R=$PWD/oracle/_ref/libwdsp_ref.so; G=$PWD/quisk_b200/libquisk_cuda.so
A="48000 3 2000 100000 1000 4 100"
QUISK_WDSP_LIB=$R python tests/quisk_swapin_driver.py $F/ref /tmp/r1.npz $A | tail -1
QUISK_WDSP_LIB=$R python tests/quisk_swapin_driver.py $F/ref /tmp/r2.npz $A | tail -1
QUISK_SWAPIN_ULP=5 QUISK_WDSP_LIB=$R python tests/quisk_swapin_driver.py $F/ref /tmp/r3.npz $A | tail -1
QUISK_WDSP_LIB=$G python tests/quisk_swapin_driver.py $F/cuda /tmp/g1.npz $A | tail -1
QUISK_WDSP_LIB=$G python tests/quisk_swapin_driver.py $F/cuda /tmp/g2.npz $A | tail -1
QUISK_SWAPIN_ULP=5 QUISK_WDSP_LIB=$G python tests/quisk_swapin_driver.py $F/cuda /tmp/g3.npz $A | tail -1
python - <<'PY'
import numpy as np
from oracle import quisk_oracle as O
L=lambda n: np.load('/tmp/%s.npz'%n)['audio']
for a,b in (('r1','r2'),('r1','r3'),('g1','g2'),('g1','g3'),('r1','g1'),('r3','g3')):
    x,y=L(a),L(b); d=np.abs(x-y); nz=np.nonzero(d>1e-9*np.abs(x).max())[0]
    print(a,b,O.rel_rms(x,y), d.max()/np.abs(x).max(), nz[:1], len(nz))
PY
