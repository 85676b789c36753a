"""(f-4) Can a device-side size ESTIMATE of a filtered stream stand in for deflate when the evaluator ranks
candidates, so that only the best stream crosses PCIe?  Offline experiment on the CPU (oracle streams + zlib level 7, the
evaluator's libdeflater level, evaluate.rs / lib.rs:391-402): order-0 entropy, order-1 (bigram-context) entropy and an
MDL-style bigram cost against zlib's choice on 58 reference fixtures + 12 synthetic pictures.

Result (profiles/r02_f4_estimator_experiment.txt): the best estimator picks zlib's winner for ~60 % of the images with a
mean size regret of 3.2-3.9 %; order-0 entropy for 27-36 %.  Not good enough to drop streams without changing the output:
the exact deflate stays on the host and all K filtered streams keep crossing PCIe (the e2e copy ceiling in bench.py).

    python profiles/r02_f4_estimator_experiment.py
"""
import sys, zlib, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, oracle
oracle.use_native()
golden = helpers.load_golden()
def H0(b):
    c=np.bincount(b,minlength=256).astype(np.float64); n=c.sum(); c=c[c>0]
    return float((c*np.log2(n/c)).sum()/8)
def H1(b):
    if b.size<2: return H0(b)
    k=(b[:-1].astype(np.int64)<<8)|b[1:]
    c=np.bincount(k,minlength=65536).astype(np.float64).reshape(256,256)
    ca=c.sum(1,keepdims=True)
    with np.errstate(divide='ignore',invalid='ignore'):
        t=np.where(c>0,c*np.log2(ca/c),0.0)
    return float(t.sum()/8)+1
def bigcost(b):  # MDL-ish: order-1 entropy + model cost of distinct bigrams
    if b.size<2: return H0(b)
    k=(b[:-1].astype(np.int64)<<8)|b[1:]
    c=np.bincount(k,minlength=65536)
    d=(c>0).sum()
    return H1(b)+d*1.0
sets={'o2':[("None",oracle.BASIC,0),("Sub",oracle.BASIC,1),("Entropy",oracle.ENTROPY,0),("Bigrams",oracle.BIGRAMS,0)],
      'c5':[("None",oracle.BASIC,0),("Bigrams",oracle.BIGRAMS,0),("BigEnt",oracle.BIGENT,0)]}
cases=[(c.name,c.ihdr,c.data) for c in golden if c.data.size>2000]
for i in range(6):
    ih,d=helpers.synth_image(512,256,6,8,100+i,"photo"); cases.append((f"photo{i}",ih,d))
    ih,d=helpers.synth_image(400,300,2,8,200+i,"photo"); cases.append((f"rgb{i}",ih,d))
for sname,S in sets.items():
    agree={k:0 for k in ('H0','H1','mix','big')}; regret={k:0.0 for k in agree}; n=0
    for name,ih,d in cases:
        outs=[oracle.filter_image(ih,d,k,b)[0] for _,k,b in S]
        z=[len(zlib.compress(o.tobytes(),7)) for o in outs]
        ests={'H0':[H0(o) for o in outs],'H1':[H1(o) for o in outs]}
        ests['mix']=[min(a,b) for a,b in zip(ests['H0'],ests['H1'])]
        ests['big']=[bigcost(o) for o in outs]
        zb=int(np.argmin(z)); n+=1
        for k,e in ests.items():
            eb=int(np.argmin(e)); agree[k]+= (z[eb]==z[zb]); regret[k]+= (z[eb]-z[zb])/z[zb]
    print(sname,n,{k:f"{100*agree[k]/n:.0f}% agree, mean regret {100*regret[k]/n:.2f}%" for k in agree})
