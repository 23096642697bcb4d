import sys, torch, time
sys.path.insert(0, ".")
import gymwipe_b200
n=65536
for mode in ("reference","mask_philox"):
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, mode=mode, strict=False); env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    T=264
    dev = torch.randint(0,2,(T,n),generator=g,device="cuda",dtype=torch.int32); dur = torch.randint(0,20,(T,n),generator=g,device="cuda",dtype=torch.int32)
    for t in range(8): env.step({"device":dev[t],"duration":dur[t]})
    torch.cuda.synchronize()
    for lo,hi,name in ((8,88,"productive"),(200,264,"degenerate")):
        if name=="degenerate":
            for t in range(88,200): env.step({"device":dev[t],"duration":dur[t]})
            torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(lo,hi): env.step({"device":dev[t],"duration":dur[t]})
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/(hi-lo)
        print(mode, name, "ms/step %.4f"%ms, "env-steps/s %.3e"%(n/ms*1e3))
