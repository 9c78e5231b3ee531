"""fwd2p (C3): half-warp = bin row, 2 chunks of 16 sample columns, 8-byte taps served per half-warp."""
import numpy as np, sys
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from mobulaop_b200 import workloads as WL
wl = WL.CONFIGS[sys.argv[1] if len(sys.argv)>1 else 'C3']
mode = sys.argv[2] if len(sys.argv)>2 else 'interleaved'   # interleaved: h -> rows 2it+h ; blocked: h -> rows h*ITERS+it
rois = WL.make_rois(wl, 0)[:2000]
H,W = wl.H, wl.W
PH,PW = wl.pooled
f32=np.float32
def axis(start, binsz, P, extent):
    lo=[]
    for p in range(P):
        for i in range(2):
            x = float(f32(start + f32(p)*binsz) + f32(f32(i+0.5)*binsz/f32(2)))
            if x < -1.0 or x > extent: lo.append(extent+1); continue
            if x<=0: x=0.0
            l=int(x)
            if l>=extent-1: l=extent-1
            lo.append(l)
    return lo
def wf(xs):
    banks={}
    for a in set(xs): banks.setdefault(a%16,set()).add(a)
    return max(len(v) for v in banks.values()) if xs else 0
ITERS=(PH+1)//2
base=0; new=0
for r in rois:
    x1,y1,x2,y2=[f32(v)*f32(wl.scale) for v in r[1:]]
    rw=max(x2-x1,f32(1)); rh=max(y2-y1,f32(1))
    ylo=axis(y1, f32(rh/f32(PH)), PH, H); xlo=axis(x1, f32(rw/f32(PW)), PW, W)
    nch=(2*PW+15)//16
    for h in range(2):
        st=None
        for it in range(ITERS):
            ph = 2*it+h if mode=='interleaved' else h*ITERS+it
            if ph>=PH: ph=PH-1
            for iy in range(2):
                y=ylo[2*ph+iy]
                need=(0,1)
                if st is not None and y==st: need=()
                elif st is not None and y==st+1: need=(1,)
                if mode=='interleaved' and iy==0: need=(0,1)
                for ch in range(nch):
                    cols=[xlo[min(16*ch+q,2*PW-1)] for q in range(16)]
                    for dx in (0,1):
                        w=wf([c+dx for c in cols])
                        base+=2*w
                        new+=len(need)*w
                st=y
print(wl.name, mode, 'wavefronts per (ROI, channel pair): base %.1f reuse %.1f'%(base/len(rois), new/len(rois)))
