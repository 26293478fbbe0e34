"""CPU copy rate into / out of pinned (cudaHostAlloc) memory against pageable memory, one thread and many (the byte surface
stages every operand frame with a memcpy into a pinned lane buffer)."""
import threading, time
import numpy as np
import torch

def rate(dst, src, reps=5):
    dst[:] = src
    t = time.perf_counter()
    for _ in range(reps):
        dst[:] = src
    return reps * src.nbytes / (time.perf_counter() - t) / 1e9

n = 256 << 20
src = np.ones(n, dtype=np.uint8)
page = np.empty(n, dtype=np.uint8)
pin = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
print("1 thread  pageable -> pageable GB/s", round(rate(page, src), 1))
print("1 thread  pageable -> pinned   GB/s", round(rate(pin, src), 1))
print("1 thread  pinned   -> pageable GB/s", round(rate(page, pin), 1))
small_src = [np.ones(88 << 10, dtype=np.uint8) for _ in range(512)]
t = time.perf_counter()
for k in range(2048):
    pin[(k % 1024) * (136 << 10):(k % 1024) * (136 << 10) + (88 << 10)] = small_src[k % 512]
print("1 thread  88 KB pieces -> pinned: us per piece", round((time.perf_counter() - t) / 2048 * 1e6, 1))
for nt in (4, 16, 32):
    chunk = n // nt
    def w(i):
        for _ in range(4):
            pin[i * chunk:(i + 1) * chunk] = src[i * chunk:(i + 1) * chunk]
    ts = [threading.Thread(target=w, args=(i,)) for i in range(nt)]
    t = time.perf_counter(); [x.start() for x in ts]; [x.join() for x in ts]
    print(nt, "threads pageable -> pinned GB/s", round(4 * n / (time.perf_counter() - t) / 1e9, 1), flush=True)
