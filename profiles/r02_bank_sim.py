"""Replays the bigram scorer's lane -> window map of k_fast on the c5 bench images and counts shared-memory wavefronts
per ATOMS (= max over banks of the number of distinct counters the 32 lanes of one instruction address in that bank;
lanes on the same counter are merged by ATOMS.POPC.INC).  Run on the CPU:  python profiles/r02_bank_sim.py [noise]

xor2 + rot = the round-1 layout (two lane-selected replicas, bank = first ^ second, windows rotated by lane & 3):
3.06, the 3.1 ncu measured.  add1 without rotation = the round-2 layout (one replica, bank = second + 5 * first): 1.85.
"""
import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from oxipng_b200 import synth
W=H=1024; ch=4
img = synth.photo_image(W,H,ch,8,0xB2000001,'cpu').numpy().reshape(H,W*ch).astype(np.int32)
noise = int(sys.argv[1]) if len(sys.argv)>1 else 0
if noise:
    rng=np.random.default_rng(1); nz=rng.integers(-noise,noise+1,img.shape); nz[:,3::4]=0; img=(img+nz)&255
def paeth(a,b,c):
    p=a+b-c; pa=abs(p-a); pb=abs(p-b); pc=abs(p-c)
    return np.where((pa<=pb)&(pa<=pc),a,np.where(pb<=pc,b,c))
rows=range(200,232)
res={}
def wavefronts(addr_words):  # addr_words: [n_instr,32] word addresses
    bank=addr_words&31
    wf=np.zeros(addr_words.shape[0],dtype=np.int64)
    for b in range(32):
        m=(bank==b)
        # distinct addresses in bank b per instr
        a=np.where(m,addr_words,-1)
        a=np.sort(a,axis=1)
        d=(a[:,1:]!=a[:,:-1])&(a[:,1:]>=0)
        cnt=d.sum(1)+(a[:,0]>=0)
        wf=np.maximum(wf,cnt)
    return wf
schemes={}
def addr(scheme,f,rep,first,second):
    # returns word index
    if scheme=='xor2': return (first<<8)|((f-1)<<6)|(rep<<5)|((second^first^(rep<<4))&31)
    if scheme=='xor1': return (first<<7)|((f-1)<<5)|((second^first)&31)
    if scheme=='add1': return (first<<7)|((f-1)<<5)|((second+5*first)&31)
    if scheme=='add2': return (first<<8)|((f-1)<<6)|(rep<<5)|(((second+5*first)^(rep<<4))&31)
    if scheme=='add1b': return (first<<7)|((f-1)<<5)|((second+11*first)&31)
    if scheme=='add1f': return (first<<7)|((f-1)<<5)|((second+5*first+8*(f-1))&31)
    if scheme=='plain1': return (first<<7)|((f-1)<<5)|(second&31)
tot={s:[0,0] for s in ['xor2','xor1','add1','add2','add1b','add1f','plain1']}
totn={s:[0,0] for s in tot}
for r in rows:
    x=img[r]; u=img[r-1]
    l=np.concatenate([np.zeros(4,np.int32),x[:-4]]); ul=np.concatenate([np.zeros(4,np.int32),u[:-4]])
    preds=[l,u,(l+u)>>1,paeth(l,u,ul)]
    for fi,p in enumerate(preds):
        f=fi+1
        t=(x-p+16)&255
        tp=np.concatenate([[f+16],t[:-1]])
        # words: 1024 words; thread owns words 2p,2p+1; warp it: 8 warps x 2 iterations... thread tid handles pair pr=tid+256*it
        for rot in (True,False):
          T=t.reshape(-1,4); TP=tp.reshape(-1,4)   # [word, byte]
          nwords=T.shape[0]
          pair=np.arange(nwords)//2; lane=pair%32; warpit=pair//32  # instruction group
          i=np.arange(nwords)%2
          for j in range(4):
            sel=(j+(lane&3))&3 if rot else np.full(nwords,j)
            first=TP[np.arange(nwords),sel]; second=T[np.arange(nwords),sel]
            ok=(first<32)&(second<32)
            rep=(lane>>2)&1
            for s in tot:
                a=addr(s,f,rep,first,second)
                # group per (warpit,i): 32 lanes
                A=np.full((nwords//64,2,32),-1,dtype=np.int64)
                A[warpit,i,lane]=np.where(ok,a,-1)
                A=A.reshape(-1,32)
                # mask -1: treat as no access
                bank=A&31
                wf=np.zeros(A.shape[0],dtype=np.int64)
                for b in range(32):
                    aa=np.where((bank==b)&(A>=0),A,-1); aa=np.sort(aa,axis=1)
                    d=(aa[:,1:]!=aa[:,:-1])&(aa[:,1:]>=0)
                    wf=np.maximum(wf,d.sum(1)+(aa[:,0]>=0))
                k=0 if rot else 1
                tot[s][k]+=wf.sum(); totn[s][k]+=len(wf)
for s in tot:
    print(s,'rot: %.2f  norot: %.2f'%(tot[s][0]/totn[s][0],tot[s][1]/totn[s][1]))
