"""H2D bandwidth of 64 MiB copies from pinned host memory: default vs write-combined allocation (not shipped)."""
import ctypes as C, time, torch
rt = C.CDLL("libcudart.so")
torch.cuda.init(); torch.zeros(1, device="cuda")
N = 64 << 20
d = torch.empty(N, dtype=torch.uint8, device="cuda")
for name, flags in (("default", 0), ("portable", 1), ("write-combined", 4)):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), C.c_uint(flags)) == 0
    C.memset(p, 1, N)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), p, C.c_size_t(N), C.c_int(1), C.c_void_p(st))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-16s %.1f GB/s" % (name, 20 * N / dt / 1e9))
    rt.cudaFreeHost(p)
h = torch.empty(N, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print("torch pin_memory  %.1f GB/s" % (20 * N / (time.perf_counter() - t0) / 1e9))
