import sys; sys.path.insert(0,'/root/repo')
import torch, lightprop2d_b200 as lp
def ev(fn, it=5):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/it
n=nb=512
for M,Bt in ((256,1),(256,64)):
    modes=torch.randn(M,nb*nb,dtype=torch.complex64,device='cuda')
    fld=torch.randn(Bt,n,n,dtype=torch.complex64,device='cuda')
    b=lp.Beam2D(400e-4,n,632e-7,init_field=fld,complex_bits=64,use_gpu=True,numpy_output=False)
    c=b.deconstruct_by_modes(modes)
    t=ev(lambda: b.construct_by_modes(modes,c))
    print(f"construct_by_modes M={M} Bt={Bt}: {t:.3f} ms (includes fft2? no: lazy)")
