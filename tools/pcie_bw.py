import torch, time
d = torch.empty(5*1024*1024//4, device='cuda'); h = torch.empty(5*1024*1024//4).pin_memory()
for n in (0.5, 1.5, 5):
    m = int(n*1024*1024//4)
    for _ in range(5): h[:m].copy_(d[:m], non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(200): h[:m].copy_(d[:m], non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/200
    print(f'D2H {n} MB: {dt*1e6:.1f} us  {n/1024/dt:.1f} GB/s')
    for _ in range(5): d[:m].copy_(h[:m], non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(200): d[:m].copy_(h[:m], non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/200
    print(f'H2D {n} MB: {dt*1e6:.1f} us  {n/1024/dt:.1f} GB/s')
# sync latency
t0=time.perf_counter()
for _ in range(200):
    h[:16].copy_(d[:16], non_blocking=True); torch.cuda.current_stream().synchronize()
print('tiny copy+sync us', (time.perf_counter()-t0)/200*1e6)
