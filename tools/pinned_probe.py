"""H2D bandwidth from differently allocated pinned host buffers (tools only)."""
import ctypes, mmap, os, sys
import numpy as np, torch
dev = torch.device('cuda:0')
NB = 25165824
dst = torch.empty(NB, dtype=torch.uint8, device=dev)
libc = ctypes.CDLL('libc.so.6', use_errno=True)
def timed(src, n=20):
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dst.copy_(src, non_blocking=True); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.max(ts))
def huge_pinned(nbytes):
    sz = (nbytes + (2 << 20) - 1) & ~((2 << 20) - 1)
    m = mmap.mmap(-1, sz + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    buf = (ctypes.c_char * (sz + (2 << 20))).from_buffer(m)
    base = ctypes.addressof(buf)
    al = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
    rc = libc.madvise(ctypes.c_void_p(al), ctypes.c_size_t(sz), 14)   # MADV_HUGEPAGE
    t = torch.frombuffer(m, dtype=torch.uint8, count=nbytes, offset=al - base)
    t.zero_()                      # touch: fault the pages in (as huge pages if THP allows)
    r = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)
    return t, m, rc, r
print('THP:', open('/sys/kernel/mm/transparent_hugepage/enabled').read().strip())
keep = []
for i in range(4):
    a = torch.empty(NB, dtype=torch.uint8).pin_memory(); keep.append(a)
    print('torch pin_memory   #%d  median %.3f max %.3f ms' % (i, *timed(a)))
for i in range(4):
    t, m, rc, r = huge_pinned(NB); keep.append((t, m))
    print('hugepage+register  #%d  madvise rc %d reg %s median %.3f max %.3f ms' % (i, rc, r, *timed(t)))
print('AnonHugePages:', [l for l in open('/proc/meminfo') if 'AnonHuge' in l])
