"""Simulates shared-memory wavefronts of resident_fwd2r (two ROIs per warp, half-warp = ROI, lane = sample column)
with and without row reuse."""
import numpy as np, sys
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from mobulaop_b200 import workloads as WL
wl = WL.CONFIGS[sys.argv[1] if len(sys.argv)>1 else 'C2']
rois = WL.make_rois(wl, 0)
H,W = wl.H, wl.W
PH,PW = wl.pooled
pitch = (W+6)//4*4
if (pitch//4)%2==0: pitch+=4

f32=np.float32
def axis(start, binsz, P, extent):
    lo=[];fr=[]
    for p in range(P):
        for i in range(2):
            x = f32(start + f32(p)*binsz) + f32(f32(i+0.5)*binsz/f32(2))
            x=float(x)
            if x < -1.0 or x > extent: lo.append(extent+1); continue
            if x<=0: x=0.0
            l=int(x)
            if l>=extent-1: l=extent-1
            lo.append(l)
    return lo
def wavefronts(addr_lists):
    # addr_lists: list of word addresses accessed by the active lanes (one LDS.32 instr)
    if not addr_lists: return 0
    banks={}
    for a in set(addr_lists):
        banks.setdefault(a%32,set()).add(a)
    return max(len(v) for v in banks.values())
tot_base=0; tot_new=0; n_instr_base=0; n_lds_new=0
geo=[]
for r in rois:
    x1,y1,x2,y2=[f32(v)*f32(wl.scale) for v in r[1:]]
    rw=max(x2-x1,f32(1)); rh=max(y2-y1,f32(1))
    ylo=axis(y1, f32(rh/f32(PH)), PH, H); xlo=axis(x1, f32(rw/f32(PW)), PW, W)
    geo.append((ylo,xlo))
for j in range(0,len(geo),2):
    A=geo[j]; B=geo[min(j+1,len(geo)-1)]
    # base: per step (ph), 8 LDS: rows ylo(iy0), ylo(iy1) x {row, row+1} x {x, x+1}
    stA=None; stB=None
    for ph in range(PH):
        for iy in range(2):
            k=2*ph+iy
            rows=[]
            for (ylo,xlo),st in ((A,stA),(B,stB)):
                rows.append(ylo[k])
            # base: 4 LDS for this sample row: (row, x), (row, x+1), (row+1, x), (row+1,x+1)
            for dy in (0,1):
                for dx in (0,1):
                    addrs=[]
                    for (ylo,xlo) in (A,B):
                        for x in xlo: addrs.append((ylo[k]+dy)*pitch + x+dx)
                    tot_base+=wavefronts(addrs); n_instr_base+=1
            # new: reuse
            need=[]
            for idx,((ylo,xlo),st) in enumerate(((A,stA),(B,stB))):
                y=ylo[k]
                if st is not None and y==st: need.append(())
                elif st is not None and y==st+1: need.append((1,))
                else: need.append((0,1))
            for dy in (0,1):
                for dx in (0,1):
                    addrs=[]
                    for idx,(ylo,xlo) in enumerate((A,B)):
                        if dy in need[idx]:
                            for x in xlo: addrs.append((ylo[k]+dy)*pitch + x+dx)
                    wv=wavefronts(addrs)
                    tot_new+=wv; n_lds_new+= 1 if wv else 0
            stA=A[0][k]; stB=B[0][k]
print(wl.name,'pitch',pitch,'ROIs',len(geo))
print('base: wavefronts per ROI %.1f (instr %.1f) -> %.2f per instr'%(tot_base/len(geo), n_instr_base/len(geo)*1.0, tot_base/n_instr_base))
print('reuse: wavefronts per ROI %.1f, non-empty LDS per ROI %.1f'%(tot_new/len(geo), n_lds_new/len(geo)))
