"""How long does cudaHostRegister take on a touched 3 GB buffer (4 KiB pages vs transparent huge pages), whole and in
256 MB pieces?  Decides whether the file pipeline should pin its contig buffer instead of staging through the ring."""
import ctypes, mmap, time, traceback
import torch
rt = torch.cuda.cudart()
torch.zeros(1, device="cuda:0")
libc = ctypes.CDLL("libc.so.6", use_errno=True)
N = 3 << 30
step = 256 << 20
for huge in (False, True):
    try:
        m = mmap.mmap(-1, N + (2 << 20))
        addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
        base = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
        if huge:
            print("madvise rc", libc.madvise(ctypes.c_void_p(base), ctypes.c_size_t(N), 14), flush=True)     # MADV_HUGEPAGE
        t0 = time.perf_counter()
        ctypes.memset(base, 65, N)
        t1 = time.perf_counter()
        r = rt.cudaHostRegister(base, N, 0)
        t2 = time.perf_counter()
        print(f"huge={huge}: touch {t1 - t0:.3f} s, register 3 GB {t2 - t1:.3f} s (rc {r})", flush=True)
        r = rt.cudaHostUnregister(base)
        t3 = time.perf_counter()
        print(f"   unregister {t3 - t2:.3f} s (rc {r})", flush=True)
        ts = []
        for off in range(0, N, step):
            a = time.perf_counter(); r = rt.cudaHostRegister(base + off, step, 0); ts.append(time.perf_counter() - a)
        print(f"   12 x 256 MB {sum(ts):.3f} s (max {max(ts) * 1e3:.1f} ms, last rc {r})", flush=True)
        for off in range(0, N, step):
            rt.cudaHostUnregister(base + off)
    except Exception:
        traceback.print_exc()
