import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, oxipng_b200 as ox
from oxipng_b200 import synth
for (w,h,ch,bd,n) in [(4096,4096,4,8,'c1'),(1024,1024,4,8,'1k')]:
    d = synth.photo_image(w,h,ch,bd,1,torch.device('cuda:0')).cpu().numpy()
    hdr = ox.IhdrData(w,h,ox.ColorType.RGBA(),ox.BitDepth(bd))
    out,_ = ox.PngImage(hdr,d).filter_image(ox.FilterStrategy.MinSum, False)
    for i in range(3):
        t=time.perf_counter(); r=ox.unfilter_image(hdr,out); dt=time.perf_counter()-t
    print(n, 'unfilter host-call ms', round(dt*1e3,2), 'GB/s', round(d.size/dt/1e9,2), np.array_equal(r,d))
    for i in range(3):
        t=time.perf_counter(); o2,_=ox.PngImage(hdr,d).filter_image(ox.FilterStrategy.MinSum, False); dt=time.perf_counter()-t
    print(n, 'filter(MinSum) host-call ms', round(dt*1e3,2), 'GB/s', round(d.size/dt/1e9,2))
