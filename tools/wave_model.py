"""How should S slices be cut into waves?  Tile-count model of one forward: a convolution layer with `tpl` tiles per slice runs
ceil(b * tpl / 148) rounds of CTAs on a b-slice wave (persistent CTAs, one per SM), each round costing that layer's measured
time per round (tools/conv_probe log of one 37-slice wave).  Dynamic programme over the wave sizes; `ovh` = a fixed cost per wave
(launch ramps of its 26 kernels).  Result: without a per-wave cost the engine's split - waves of 37 and a
short tail - is optimal for every S; 74-slice waves only save the per-wave cost (0.65 % at 0.1 ms per wave; measured: none).

    python tools/wave_model.py [profiles/r02_call10_single.log]
"""
import sys
LOG = sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_call10_single.log"
import math
import re
layers=[]
for l in open(LOG):
    if l.startswith('TIME'):
        m=re.search(r'TIME\s+(\S+)\s+N=37\s+(\d+)x\d+\s+C=\s*(\d+)\+(\d+)\s*->\s*(\d+) taps=(\d): ([\d.]+) ms',l)
        name,hw,c0,c1,co,taps,ms=m.groups(); hw=int(hw);co=int(co);ms=float(ms)
        bn=64 if co==64 else 128
        tpl=hw*hw//128*(co//bn)
        layers.append((name,tpl,ms))
def cost(b,ovh=0.0):
    c=0
    for name,tpl,ms in layers:
        tau=ms/math.ceil(37*tpl/148)
        c+=math.ceil(b*tpl/148)*tau
    return c+ovh
w37=cost(37)
print('wave37',w37)
for b in (4,9,18,27,37,41,46,50,55,64,74): print(b, round(cost(b),3), round(cost(b)/b/(w37/37),4))
def plan(S,cap,ovh):
    best=[0]+[1e9]*S; ch=[0]*(S+1)
    for s in range(1,S+1):
        for b in range(1,min(cap,s)+1):
            v=best[s-b]+cost(b,ovh)
            if v<best[s]-1e-12: best[s]=v; ch[s]=b
    out=[];s=S
    while s: out.append(ch[s]); s-=ch[s]
    return best[S],sorted(out,reverse=True)
for S in (300,512,100,150,600):
    for cap in (37,74):
        for ovh in (0.0,0.1):
            t,p=plan(S,cap,ovh)
            cur=sum(cost(min(37,S-s0),ovh) for s0 in range(0,S,37))
            print(S,cap,ovh,round(t,3),'current',round(cur,3),'gain %.2f%%'%(100*(cur-t)/cur),p)
