"""GPU micro-benchmark of the tcgen05 GEMM core on the denoiser's shapes (config 3: M = 21312)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmd_b200 import _lib

def bench(M, N, K, aux=False, act=0, out_f32=False, mode=0, reps=20, cublas=True):
    dev = 'cuda'
    dt = torch.bfloat16 if mode == 0 else torch.float32
    x = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    xl = torch.zeros_like(x) if mode else None; wl = torch.zeros_like(w) if mode else None
    b = torch.randn(N, device=dev)
    a = torch.randn(M, N, device=dev).to(torch.float32 if mode else torch.bfloat16) if aux else None
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    def run():
        _lib.check(_lib.lib().msmd_linear(mode, p(x), p(xl), p(w), p(wl), p(b), p(a), p(out), M, N, K, K, K, N, N if aux else 0,
                                          int(out_f32), int(mode == 1), act, _lib.stream_ptr()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * M * N * K / ms / 1e9
    print(f'mode{mode} M={M} N={N} K={K} aux={int(aux)} act={act} f32out={int(out_f32)}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s')
    # cuBLAS comparison (library baseline)
    if mode == 0 and cublas:
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3): torch.matmul(x, w.t(), out=y)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps): torch.matmul(x, w.t(), out=y)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / reps
        print(f'      cuBLAS bf16 (no epilogue): {ms2*1e3:8.1f} us  {2.0*M*N*K/ms2/1e9:7.1f} TFLOP/s')

if __name__ == '__main__':
    M = 21312
    bench(M, 512, 512)
    bench(M, 512, 2048)
    bench(M, 768, 3072)
    bench(M, 1536, 512)
    bench(M, 512, 512, aux=True, out_f32=True)
    bench(M, 2048, 512, act=1)
    bench(M, 512, 2048, aux=True, out_f32=True)
    bench(M, 256, 512, act=1)
    bench(M, 80, 256, out_f32=True)
    bench(8192, 8192, 8192)
    bench(M, 512, 512, aux=True, out_f32=True, mode=1)
    bench(8192, 15069 // 3 * 3 // 4 * 4, 448, out_f32=True, mode=1)
