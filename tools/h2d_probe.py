import torch, time, os, subprocess
dev=torch.device('cuda',0)
print(subprocess.run("lscpu | grep -i 'numa\\|^CPU(s)\\|Model name'; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c", shell=True, capture_output=True, text=True).stdout)
W,N=65536,32
gen=torch.Generator(device=dev); gen.manual_seed(1)
ring=[torch.randint(0,5,(W,N),generator=gen,device=dev,dtype=torch.int8) for _ in range(8)]
def timeit(bufs, label):
    d=torch.empty((W,N),dtype=torch.int8,device=dev)
    res=[]
    for b in bufs:
        for _ in range(3): d.copy_(b, non_blocking=True)
        torch.cuda.synchronize()
        a=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): d.copy_(b, non_blocking=True)
        e.record(); torch.cuda.synchronize()
        res.append(round(a.elapsed_time(e)/10*1e3,1))
    print(label, res, "us per 2 MB H2D")
timeit([r.cpu().pin_memory() for r in ring], "separate .cpu().pin_memory():")
big=torch.empty((8,W,N),dtype=torch.int8).pin_memory()
for i,r in enumerate(ring): big[i].copy_(r)
timeit([big[i] for i in range(8)], "slices of one pinned allocation:")
big2=torch.empty((8,W,N),dtype=torch.int8,pin_memory=True)
timeit([big2[i] for i in range(8)], "slices of torch.empty(pin_memory=True):")
timeit([torch.randint(0,5,(W,N),dtype=torch.int8).pin_memory() for _ in range(8)], "cpu randint .pin_memory():")
