"""Where the API e2e step spends its time on a B200 (host legs vs chain), printed as a table: a tuning aid."""
import sys, time
sys.path.insert(0, '.')
import torch
from lzy_b200 import _native as N
from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op
shape=(16384,16384)
@op
def ident(t: torch.Tensor) -> torch.Tensor:
    return t
rt=B200Runtime(gpus=[0], placement="single"); lzy=Lzy(runtime=rt); rt._ensure_ready()
x=torch.empty(shape,device="cuda:0").normal_()
host_in=torch.empty(shape,pin_memory=True).copy_(x); host_out=torch.empty(shape,pin_memory=True)
dev_in=torch.empty_like(x)
def sync(): torch.cuda.synchronize()
def t(fn, n=5):
    fn(); sync(); t0=time.perf_counter()
    for _ in range(n): fn()
    sync(); return (time.perf_counter()-t0)/n*1e3
def step(inp):
    with lzy.workflow("p"):
        y=inp
        for _ in range(8): y=ident(y)
        out=materialize(y)
    rt.clear(); return out
print("h2d torch copy_ into dev tensor      %.2f ms" % t(lambda: dev_in.copy_(host_in, non_blocking=True)))
print("h2d x.to(cuda)                        %.2f ms" % t(lambda: host_in.to("cuda:0", non_blocking=True)))
s=N.SlotHandle.create_output("probe", 0, x.numel()*4)
def ph():
    global s
    s.close(); s=N.SlotHandle.create_output("probe", 0, x.numel()*4); s.pack_host(host_in.data_ptr(), x.numel()*4, int(torch.cuda.current_stream().cuda_stream)); s.commit()
print("h2d lzyb pack_host + commit          %.2f ms" % t(ph))
print("d2h host_out.copy_(x)                 %.2f ms" % t(lambda: host_out.copy_(x, non_blocking=True)))
print("chain step(x)                         %.2f ms" % t(lambda: step(x)))
print("chain step(dev_in)                    %.2f ms" % t(lambda: step(dev_in)))
def full():
    dev_in.copy_(host_in, non_blocking=True); y=step(dev_in); host_out.copy_(y, non_blocking=True)
print("full e2e                              %.2f ms" % t(full))
def full_sync():
    dev_in.copy_(host_in, non_blocking=True); sync(); y=step(dev_in); sync(); host_out.copy_(y, non_blocking=True)
print("full e2e with syncs between legs      %.2f ms" % t(full_sync))
# the unsynchronised loop again, with events between its legs: where on the GPU timeline the step's time goes
def timeline(n=4):
    evs=[]
    for _ in range(n):
        e=[torch.cuda.Event(enable_timing=True) for _ in range(4)]; h=[]
        e[0].record(); h.append(time.perf_counter()); dev_in.copy_(host_in, non_blocking=True); e[1].record(); h.append(time.perf_counter())
        y=step(dev_in); e[2].record(); h.append(time.perf_counter()); host_out.copy_(y, non_blocking=True); e[3].record(); h.append(time.perf_counter()); evs.append((e,h))
    sync()
    for i,(e,h) in enumerate(evs):
        gap = evs[i-1][0][3].elapsed_time(e[0]) if i else 0.0
        print("  iter %d gpu: gap %.2f h2d %.2f chain %.2f d2h %.2f | host: h2d %.2f step %.2f d2h %.2f ms" % (i, gap, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), (h[1]-h[0])*1e3, (h[2]-h[1])*1e3, (h[3]-h[2])*1e3))
full(); sync(); timeline()
rt.close()
