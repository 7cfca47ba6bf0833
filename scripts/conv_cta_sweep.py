import torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cutie_b200.kernels as K_
torch.backends.cudnn.allow_tf32=False
def timed(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e3
Cin,Cout=512,128
w=torch.randn(Cout,Cin,3,3,device='cuda')*0.02
img=K_.conv_weight_image(w)
b=torch.randn(Cout,device='cuda')
print('3x3 512->128, one 2x54 tile (N=112) per CTA, 144 steps each, direct path: CTAs -> us')
for n in [64,96,112,120,124,128,130,132,134,136,138,140,142,144,146,148,152,160]:
    x=torch.randn(1,Cin,2*n,54,device='cuda').contiguous(memory_format=torch.channels_last)   # n tiles of 2 rows
    t=timed(lambda: K_.conv_tc(x,img,b,Cout,ksize=3,units_per_cta=Cin//32))
    print(n, round(t,1))
