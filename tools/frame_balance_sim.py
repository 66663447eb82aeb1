"""Offline study of where the frames of a visibility launch should run (developer tool).

    VX_B200_FRAME_STATS=stats.txt python tools/dev_gpu_perf.py 4541 semantic_kitti 1      # on the GPU box
    python tools/frame_balance_sim.py stats.txt [alpha]

Inverts a simple contention model (a CTA among k on its SM runs at 1 / (1 + alpha (k - 1)) of its speed alone) into the
work of every frame, then replays other frame -> SM assignments.  Result (round 1): balancing the SUM of work per SM gains
3 %; the batch ends with its dearest frames, which only fewer neighbours would help (needs a good cost predictor).
"""
import numpy as np, sys
raw=np.loadtxt(sys.argv[1],dtype=np.uint64)[-909:]
a=raw.astype(float)
nep=a[:,1]; waves=a[:,4]
m=nep+1>0
sm=a[:,11].astype(int)[m]; t0=a[:,12][m]; t1=a[:,13][m]; waves=waves[m]; nep=nep[m]
live0=(raw[:,14]&np.uint64(0xFFFFFFFF)).astype(float)[m]
end=(t1-t0.min())/1e6
ALPHA=float(sys.argv[2]) if len(sys.argv)>2 else 0.066
def rate(k): return 1.0/(1.0+ALPHA*(k-1))
# invert: alone-time work per frame
W=np.zeros(len(end))
for s in range(148):
    idx=np.where(sm==s)[0]
    o=idx[np.argsort(end[idx])]
    t=0.0; acc=0.0; k=len(o)
    for j,f in enumerate(o):
        acc+= (end[f]-t)*rate(k-j); t=end[f]; W[f]=acc
print("alpha",ALPHA,"W: mean %.0f std %.0f max %.0f ; corr(W,waves) %.3f corr(W,live0) %.3f"%(W.mean(),W.std(),W.max(),np.corrcoef(W,waves)[0,1],np.corrcoef(W,live0)[0,1]))
def simulate(buckets,Wv=W):
    ends=[]
    for b in buckets:
        w=np.sort(Wv[b]); t=0.0; done=0.0; k=len(w)
        for j,x in enumerate(w):
            t+=(x-done)/rate(k-j); done=x
        ends.append(t)
    return max(ends), np.mean(ends)
# actual placement
actual=[np.where(sm==s)[0] for s in range(148)]
print("actual placement: sim end %.0f (measured %.0f) mean SM end %.0f"%(*simulate(actual)[:1],end.max(),simulate(actual)[1]))
n=len(W)
def lpt(pred,cap_extra=True):
    order=np.argsort(-pred); S=148; base=n//S
    load=np.zeros(S); b=[[] for _ in range(S)]
    for k,f in enumerate(order):
        cap=base if k<base*S else base+1
        c=[i for i in range(S) if len(b[i])<cap]
        i=min(c,key=lambda i:load[i]); b[i].append(f); load[i]+=pred[f]
    return [np.array(x) for x in b]
rng=np.random.default_rng(0)
for name,pred in (("oracle W",W),("waves",waves),("W+5% noise",W*(1+0.05*rng.standard_normal(n))),("W+10% noise",W*(1+0.10*rng.standard_normal(n))),("random",rng.random(n))):
    print("LPT sums by",name,": end %.0f mean %.0f"%simulate(lpt(pred)))
# graded: sorted consecutive groups with sizes; heavy frames fewer neighbours (needs 7-capacity launch)
def graded(pred,T):
    order=list(np.argsort(-pred)); b=[]; i=0
    while i<n:
        f=order[i]
        # largest k<=7 such that heaviest finishes by T given all k equal-ish: P*(1+alpha(k-1))<=T
        k=int(np.floor(1+(T/pred[f]-1)/ALPHA)); k=max(1,min(7,k))
        b.append(np.array(order[i:i+k])); i+=k
    return b
for name,pred in (("oracle W",W),("W+5% noise",W*(1+0.05*rng.standard_normal(n)))):
    lo,hi=pred.max(),pred.max()*2
    for _ in range(40):
        T=(lo+hi)/2
        if len(graded(pred,T))<=148: hi=T
        else: lo=T
    b=graded(pred,hi); print("graded consecutive by",name,": buckets",len(b),"sizes",np.bincount([len(x) for x in b]),"end %.0f mean %.0f"%simulate(b))
# mixed: heavy with light (sorted, bucket i gets frames i, 2S-1-i, 2S+i, ... serpentine)
def serp(pred):
    order=np.argsort(-pred); S=148; b=[[] for _ in range(S)]
    for k,f in enumerate(order):
        r,c=divmod(k,S); b[c if r%2==0 else S-1-c].append(f)
    return [np.array(x) for x in b]
print("serpentine oracle: end %.0f mean %.0f"%simulate(serp(W)))
