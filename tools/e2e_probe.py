import sys, time, torch
sys.path.insert(0, '/root/repo')
from wab_gym_b200 import VecEnv
n=4096
env=VecEnv(n, seed=0); env.reset()
hb=env.alloc_host_buffers()
acts=torch.randint(0,5,(n,),dtype=torch.uint8).pin_memory()
hb["actions"].copy_(acts)
def timeit(f, k=300):
    for _ in range(20): f()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/k*1e6
print("step_host_packed us", timeit(lambda: env.step_host(hb)))
d_act=torch.empty(n,dtype=torch.uint8,device='cuda')
def dev_only():
    d_act.copy_(hb["actions"], non_blocking=True); env.step(d_act); torch.cuda.current_stream().synchronize()
print("h2d+kernel+sync us", timeit(dev_only))
dev=torch.empty(hb["block"].numel(),dtype=torch.uint8,device='cuda')
def d2h():
    hb["block"].copy_(dev, non_blocking=True); torch.cuda.current_stream().synchronize()
print("d2h %d bytes + sync us"%dev.numel(), timeit(d2h))
half=dev.numel()//2
s2=torch.cuda.Stream()
def d2h2():
    hb["block"][:half].copy_(dev[:half], non_blocking=True)
    with torch.cuda.stream(s2): hb["block"][half:].copy_(dev[half:], non_blocking=True)
    torch.cuda.synchronize()
print("d2h split over 2 streams us", timeit(d2h2))
def kern():
    env.step(d_act); torch.cuda.current_stream().synchronize()
print("kernel+sync us", timeit(kern))
