"""GPU: the tcgen05 GEMM core (msmd_linear) against a torch fp32 reference of the same op.
bf16 mode: inputs are bf16-rounded on both sides, so the only difference is accumulation order
and the bf16 rounding of the output -> tolerance 2^-8 relative per element.
tf32x3 mode: fp32-grade, tolerance 2e-6 relative L2."""
import ctypes as C

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def run_linear(mode, x, w, bias=None, aux=None, act=0, out_f32=False):
    from msmd_b200 import _lib
    M, K = x.shape
    N = w.shape[0]
    dev = x.device
    out = torch.full((M, N), float('nan'), device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    if mode == 0:
        xa, wa = x.to(torch.bfloat16).contiguous(), w.to(torch.bfloat16).contiguous()
        xl = wl = None
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    else:
        def split(t):
            dt = torch.float32 if mode == 1 else torch.float16
            hi, lo = torch.empty_like(t, dtype=dt), torch.empty_like(t, dtype=dt)
            fn = _lib.lib().msmd_split_tf32 if mode == 1 else _lib.lib().msmd_split_f16
            _lib.check(fn(t.data_ptr(), hi.data_ptr(), lo.data_ptr(), t.numel(), _lib.stream_ptr()))
            return hi, lo
        xa, xl = split(x.float().contiguous())
        wa, wl = split(w.float().contiguous())
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    aux_f32 = int(aux is not None and aux.dtype == torch.float32)
    _lib.check(_lib.lib().msmd_linear(mode, p(xa), p(xl), p(wa), p(wl), p(bias), p(aux), p(out), M, N, K,
                                      xa.stride(0), wa.stride(0), out.stride(0),
                                      aux.stride(0) if aux is not None else 0, int(out_f32), aux_f32, act,
                                      _lib.stream_ptr()))
    torch.cuda.synchronize()
    return out


def ref_linear(mode, x, w, bias, aux, act):
    if mode == 0:
        x, w = x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double()
    else:
        x, w = x.double(), w.double()
    y = x @ w.t()
    if bias is not None:
        y = y + bias.double()
    if act:
        y = torch.nn.functional.gelu(y)
    if aux is not None:
        y = y + aux.double()
    return y


SHAPES = [(128, 256, 64), (128, 256, 512), (333, 512, 512), (21312 // 8, 1536, 512), (2664, 2048, 512),
          (2664, 512, 2048), (100, 80, 256), (4, 512, 512), (1000, 71 + 1, 256), (257, 768, 3072),
          (192, 512, 512), (500, 264, 128), (21312, 512, 512)]   # small-M 64-wide tiles; full-size CTA-pair tiles


@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_bf16_bias(built_lib, M, N, K):
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N)
    x = torch.randn(M, K, device='cuda', generator=g)
    w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    for out_f32 in (False, True):
        got = run_linear(0, x, w, b, None, 0, out_f32)
        want = ref_linear(0, x, w, b, None, 0)
        assert torch.isfinite(got.float()).all(), 'unwritten / non-finite output'
        err = ((got.double() - want).abs() / (want.abs() + 1.0)).max().item()
        assert err < (1e-4 if out_f32 else 6e-3), (M, N, K, out_f32, err)


@pytest.mark.parametrize('M,N,K', [(333, 512, 512), (2664, 2048, 512), (130, 256, 128)])
def test_gemm_bf16_gelu(built_lib, M, N, K):
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(M, K, device='cuda', generator=g)
    w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    got = run_linear(0, x, w, b, None, 1, False)
    want = ref_linear(0, x, w, b, None, 1)
    err = ((got.double() - want).abs() / (want.abs() + 1.0)).max().item()
    assert err < 6e-3, err
    got = run_linear(0, x, w, b, None, 1, True)
    err = ((got.double() - want).abs() / (want.abs() + 1.0)).max().item()
    assert err < 4e-4, err          # tanh-form GELU (fit 2.5e-5) + tanh.approx (2^-11 relative)


@pytest.mark.parametrize('M,N,K', [(333, 512, 512), (2664, 512, 2048), (64, 512, 512), (2000, 768, 768)])
@pytest.mark.parametrize('aux_dt,out_f32', [(torch.bfloat16, True), (torch.float32, True), (torch.bfloat16, False)])
def test_gemm_bf16_residual(built_lib, M, N, K, aux_dt, out_f32):
    g = torch.Generator(device='cuda').manual_seed(2)
    x = torch.randn(M, K, device='cuda', generator=g)
    w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    aux = torch.randn(M, N, device='cuda', generator=g).to(aux_dt)
    got = run_linear(0, x, w, b, aux, 0, out_f32)
    want = ref_linear(0, x, w, b, aux, 0)
    err = ((got.double() - want).abs() / (want.abs() + 1.0)).max().item()
    assert err < (1e-4 if out_f32 else 6e-3), err


@pytest.mark.parametrize('M,N,K', [(128, 128, 32), (333, 512, 512), (1000, 1672, 448), (64, 2048, 512),
                                   (777, 512, 2048)])
def test_gemm_tf32x3(built_lib, M, N, K):
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(M, K, device='cuda', generator=g)
    w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    aux = torch.randn(M, N, device='cuda', generator=g)
    got = run_linear(1, x, w, b, None, 0, True)
    # tensor-core accumulation truncates: error grows ~K/8 * 2^-25 on the hi*hi accumulator
    tol = 1.5e-6 + 2.5e-9 * K
    assert rel_l2(got, ref_linear(1, x, w, b, None, 0)) < tol
    got = run_linear(1, x, w, b, aux, 1, True)
    assert rel_l2(got, ref_linear(1, x, w, b, aux, 1)) < tol


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (333, 512, 512), (1000, 1672, 448), (64, 2048, 512),
                                   (777, 512, 2048), (2664, 1536, 512), (21312, 512, 512)])
def test_gemm_fp16x3(built_lib, M, N, K):
    """fp16 two-term split (x = hi + 2^-11 lo, 22 mantissa bits), three kind::f16 passes: fp32-grade like tf32x3,
    single-CTA and CTA-pair tiles; also exercised with operands spanning 6 orders of magnitude."""
    g = torch.Generator(device='cuda').manual_seed(6)
    x = torch.randn(M, K, device='cuda', generator=g)
    w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    aux = torch.randn(M, N, device='cuda', generator=g)
    tol = 1.5e-6 + 2.5e-9 * K
    got = run_linear(2, x, w, b, None, 0, True)
    assert rel_l2(got, ref_linear(1, x, w, b, None, 0)) < tol
    got = run_linear(2, x, w, b, aux, 1, True)
    assert rel_l2(got, ref_linear(1, x, w, b, aux, 1)) < tol
    scale = torch.logspace(-3, 3, K, device='cuda')          # columns from 1e-3 to 1e3: residuals near fp16 subnormals
    got = run_linear(2, x * scale, w / scale, b, None, 0, True)
    assert rel_l2(got, ref_linear(1, x * scale, w / scale, b, None, 0)) < 2 * tol


def test_gemm_strided_views(built_lib):
    """Row strides larger than the row (packed qkv slices) for x, out and aux."""
    g = torch.Generator(device='cuda').manual_seed(4)
    big = torch.randn(500, 1536, device='cuda', generator=g).to(torch.bfloat16)
    x = big[:, 512:1024]
    w = (torch.randn(512, 512, device='cuda', generator=g) / 22).to(torch.bfloat16)
    outbig = torch.zeros(500, 1024, device='cuda', dtype=torch.bfloat16)
    from msmd_b200 import _lib
    _lib.check(_lib.lib().msmd_linear(0, x.data_ptr(), None, w.data_ptr(), None, None, None, outbig[:, 512:].data_ptr(),
                                      500, 512, 512, 1536, 512, 1024, 0, 0, 0, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    want = x.double() @ w.double().t()
    assert ((outbig[:, 512:].double() - want).abs() / (want.abs() + 1)).max() < 6e-3
    assert (outbig[:, :512] == 0).all()
