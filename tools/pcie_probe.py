import torch, time
d=torch.empty(1<<30,dtype=torch.uint8,device='cuda'); h=torch.empty(1<<30,dtype=torch.uint8).pin_memory()
d2=torch.empty(1<<29,dtype=torch.uint8,device='cuda'); h2=torch.empty(1<<29,dtype=torch.uint8).pin_memory()
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for name,fn in [("d2h",lambda: h.copy_(d,non_blocking=True)),("h2d",lambda: d.copy_(h,non_blocking=True))]:
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print(name, 5*(1<<30)/(time.perf_counter()-t)/1e9, "GB/s")
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("duplex d2h", 5*(1<<30)/dt/1e9, "h2d", 5*(1<<29)/dt/1e9)
