import torch, time
n=100_000_000
h1=torch.empty(n,dtype=torch.uint8).pin_memory(); h2=torch.empty(n,dtype=torch.uint8).pin_memory()
d1=torch.empty(n,dtype=torch.uint8,device='cuda'); d2=torch.empty(n,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(f,reps=5):
    f(); torch.cuda.synchronize(); a=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-a)/reps*1e3
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1,non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
def both(): h2d(); d2h()
def chunks():
    k=12; m=n//k
    for i in range(k):
        with torch.cuda.stream(s1): d1[i*m:(i+1)*m].copy_(h1[i*m:(i+1)*m],non_blocking=True)
        with torch.cuda.stream(s2): h2[i*m:(i+1)*m].copy_(d2[i*m:(i+1)*m],non_blocking=True)
print('h2d 100MB ms',t(h2d),'GB/s',n/t(h2d)/1e6)
print('d2h 100MB ms',t(d2h),'GB/s',n/t(d2h)/1e6)
print('both ms',t(both))
print('chunked both ms',t(chunks))
