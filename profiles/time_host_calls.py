import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, oxipng_b200 as ox
from oxipng_b200 import synth, reduction as R
w=h=4096
d = synth.photo_image(w,h,4,8,1,torch.device('cuda:0')).cpu().numpy()
hdr = ox.IhdrData(w,h,ox.ColorType.RGBA(),ox.BitDepth(8))
img = ox.PngImage(hdr,d)
out,_ = img.filter_image(ox.FilterStrategy.MinSum, False)
for name, fn in [('unfilter_image', lambda: ox.unfilter_image(hdr,out)), ('cleaned_alpha_channel', lambda: R.cleaned_alpha_channel(img)),
                 ('interlace_image', lambda: ox.interlace_image(img)), ('reduced_to_indexed(None)', lambda: R.reduced_to_indexed(img, True))]:
    ts=[]
    for i in range(6):
        t=time.perf_counter(); r=fn(); ts.append(time.perf_counter()-t)
    print(name, 'host-call ms', [round(x*1e3,2) for x in ts])
