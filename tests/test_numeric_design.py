"""CPU: the numeric design of the tensor-core FLAME path, checked on the host.

The fused kernel (csrc/flame_tc.cu) and the pose kernel (csrc/flame.cu) feed fp16 two-term splits s*x = hi + lo to the
tensor cores, with hi and lo on ONE scale s per operand (flame.cuh: kFlameScaleA for the coefficients, kFlameScaleB for
the blendshape basis, kFlameScaleJ for the folded joint regressor) so that hi*hi, lo*hi and hi*lo can share an
accumulator.  This test restates the split in numpy and checks, on the synthetic FLAME assets the GPU tests use, that
the scales keep the residuals representable: 21+ mantissa bits for every operand of realistic magnitude, no overflow."""
import os
import re

import numpy as np

from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scales():
    src = open(os.path.join(ROOT, 'ubisoft-laforge-msmd_b200', 'csrc', 'flame.cuh')).read()
    m = re.search(r'kFlameScaleA = ([0-9.]+)f, kFlameScaleB = ([0-9.]+)f, kFlameScaleJ = ([0-9.]+)f', src)
    assert m, 'flame.cuh no longer declares the three operand scales'
    return [float(g) for g in m.groups()]


def _split(x, s):
    sx = (x.astype(np.float32) * np.float32(s)).astype(np.float32)
    hi = sx.astype(np.float16)
    lo = (sx - hi.astype(np.float32)).astype(np.float16)
    assert np.isfinite(hi.astype(np.float32)).all(), 'operand overflows fp16 at this scale'
    back = (hi.astype(np.float64) + lo.astype(np.float64)) / s
    return back


def _check(x, s, what, floor):
    back = _split(x, s)
    x64 = x.astype(np.float64)
    err = np.abs(back - x64)
    big = np.abs(x64) >= floor
    rel = (err[big] / np.abs(x64[big])).max(initial=0.0)
    print(f'{what}: scale {s:g}, worst relative error of hi + lo = {rel:.2e} over |x| >= {floor:g} '
          f'({big.mean() * 100:.1f}% of the entries), worst absolute error below = {err[~big].max(initial=0.0):.2e}')
    assert rel < 2.0 ** -20, (what, rel)                      # two fp16 terms: 11 + 11 bits minus the sign of lo
    assert err[~big].max(initial=0.0) < 2.0 ** -24 / s * 1.01, what   # subnormal residuals: absolute precision 2^-24 / s


def test_fp16_two_term_splits_keep_22_bits_on_flame_operands():
    sA, sB, sJ = _scales()
    assets = synth.flame_assets(0, synth.FLAME_V, 300, 100)
    shapedirs = np.asarray(assets['shapedirs'], dtype=np.float32)          # [V, 3, 400]
    posedirs = np.asarray(assets['posedirs'], dtype=np.float32)            # [36, V*3]
    basis = np.concatenate([shapedirs.reshape(-1, shapedirs.shape[-1]), posedirs.T], axis=1)
    _check(basis, sB, 'blendshape basis (shapedirs | posedirs)', floor=6.2e-5 * 2048 / sB)
    sh, ex, po, ey = synth.flame_inputs(512, 300, 100, seed=3)
    betas = np.concatenate([np.asarray(sh), np.asarray(ex)], axis=1).astype(np.float32)
    _check(betas, sA, 'coefficients (shape | expression)', floor=6.2e-5 * 2048 / sA)
    # folded joint regressor J_regressor @ shapedirs: [15, 400]
    Jr = np.asarray(assets['J_regressor'], dtype=np.float64)               # [5, V]
    Jb = np.einsum('jv,vck->jck', Jr, shapedirs.astype(np.float64)).reshape(15, -1).astype(np.float32)
    _check(Jb, sJ, 'folded joint regressor', floor=6.2e-5 * 2048 / sJ)
    # headroom: the largest realistic coefficient (|beta| = 10 sigma) and basis entry stay far from fp16's 65504
    assert 10.0 * sA < 6.0e4 and np.abs(basis).max() * sB < 6.0e4 and np.abs(Jb).max() * sJ < 6.0e4


def test_three_pass_fp16_embedding_matches_fp32_to_16_bit_rounding():
    """csrc/denoiser_kernels.cu embed_x_mma_kernel: feature_proj(x_t) as hi*hi + lo*hi + hi*lo over UNSCALED fp16 splits of
    x_t (the sampler state, |x| up to ~6) and of the weight (|w| <= 1/sqrt(68)).  Emulated in numpy: the three-pass sum is
    within 2e-6 of the exact product - 250x below the fp16 / 2000x below the bf16 rounding of the stored embedding."""
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((512, 67)) * np.linspace(0.05, 2.0, 512)[:, None]).astype(np.float32)
    w = rng.uniform(-1 / np.sqrt(68), 1 / np.sqrt(68), (512, 67)).astype(np.float32)
    f = lambda a: a.astype(np.float16).astype(np.float32)
    xh, wh = f(x), f(w)
    xl, wl = f(x - xh), f(w - wh)
    three = (xh.astype(np.float64) @ wh.T.astype(np.float64) + xl.astype(np.float64) @ wh.T.astype(np.float64)
             + xh.astype(np.float64) @ wl.T.astype(np.float64))
    exact = x.astype(np.float64) @ w.T.astype(np.float64)
    err = np.abs(three - exact).max()
    print(f'three-pass fp16 embedding vs exact: max abs error {err:.2e} (outputs up to {np.abs(exact).max():.2f})')
    assert err < 2e-6
