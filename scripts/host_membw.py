import numpy as np, time, threading
def run(nt, mb=256):
    srcs=[np.ones(mb<<20,dtype=np.uint8) for _ in range(nt)]; dsts=[np.empty_like(s) for s in srcs]
    for d,s in zip(dsts,srcs): d[:]=s
    def w(i):
        for _ in range(3): dsts[i][:]=srcs[i]
    ts=[threading.Thread(target=w,args=(i,)) for i in range(nt)]
    t=time.perf_counter(); [x.start() for x in ts]; [x.join() for x in ts]; dt=time.perf_counter()-t
    print(nt,"threads: copy GB/s (read+write)", round(2*3*nt*(mb<<20)/dt/1e9,1), flush=True)
for nt in (1,4,8,16): run(nt)
