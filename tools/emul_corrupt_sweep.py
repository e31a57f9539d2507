"""TEST INFRASTRUCTURE: random damaged scans (byte flips, cuts, deletions; with and without restart intervals) through the CPU
emulation of the speculative decoder (tests/emul: the kernels' own per-thread functions, units of several subsequence lengths)
against the oracle: statuses and coefficients.  No GPU needed.   python tools/emul_corrupt_sweep.py SEED CASES"""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in ('', 'video-coding_b200', 'tests'): sys.path.insert(0, os.path.join(ROOT, p_))
import numpy as np, hcjpeg, emul, synth
from oracle import pyoracle as orc
seed=int(sys.argv[1]); ncase=int(sys.argv[2])
rng=np.random.default_rng(seed)
bad=0; checked=0
for i in range(ncase):
    chroma=int(rng.choice([420,422,444])); w,h=int(rng.integers(8,200)),int(rng.integers(8,120))
    kind=int(rng.integers(3)); n=len(synth.frame(0,w,h,chroma))
    yuv = synth.frame(i,w,h,chroma) if kind==0 else bytes([int(rng.integers(256))])*n if kind==1 else rng.integers(0,256,n,dtype=np.uint8).tobytes()
    try: j=bytearray(orc.encode(yuv,w,h,chroma,int(rng.choice([20,75,95,100])),restart_interval=int(rng.choice([0,0,3,8,40]))))
    except orc.OracleError: continue
    hdr=hcjpeg.header_decode(bytes(j)).scan_byte_pos
    op=int(rng.integers(4))
    if op==0 and len(j)-hdr>6:
        for _ in range(int(rng.integers(1,6))): j[int(rng.integers(hdr,len(j)-2))]=int(rng.integers(256))
    elif op==1 and len(j)-hdr>8: j=j[:int(rng.integers(hdr+1,len(j)-2))]+b"\xff\xd9"
    elif op==2:
        p=int(rng.integers(hdr,len(j)-2)); j=j[:p]+j[p+int(rng.integers(1,6)):]
    j=bytes(j)
    try:
        dec=orc.decode(j,want_blocks=True); ost=0; want=dec.coefs_abs_dc().astype(np.int16); nb=dec.nblocks
    except orc.OracleError as e:
        ost=e.status; nb=8192
    if ost not in (0,-2,-3,-4): continue
    for T,S in ((16,128),(32,512),(64,2048)):
        st,got,_=emul.decode_units(j,nb,hdr,T=T,S=S)
        if st!=ost or (ost==0 and not np.array_equal(got,want)):
            bad+=1; print('MISMATCH case',i,(chroma,w,h),'op',op,'T,S',T,S,'emul',st,'oracle',ost); open('/tmp/sweep_bad_%d_%d.jpg'%(seed,i),'wb').write(j); break
    checked+=1
print('seed',seed,'checked',checked,'bad',bad)
