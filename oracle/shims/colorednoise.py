"""CPU restatement of `colorednoise.powerlaw_psd_gaussian` (Timmer & Koenig 1995 power-law noise).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED: the reference depends on the PyPI package `colorednoise` with an *unpinned*
version (`/root/reference/Pipfile:10`), its source is not under /root/reference and is not
installed in this image.  This file restates the published algorithm of the release that was
current when the reference was written (v1.1.1, Feb 2021); `VERSION2_SCALING` switches to the
2.x behaviour (DC and Nyquist real parts multiplied by sqrt(2)).  Single call site in the
reference: `icem/controllers/icem.py:73-75` (via `:85` and `:102`).

Draw order (this is what "identical RNG state" means for the parity tests): ONE global
`np.random` call for all real parts `sr[..., K]` (C order), then ONE for all imaginary parts.
`np.random.normal(scale=s, size)` with loc=0 equals `np.random.standard_normal(size) * s`
bit-for-bit (checked in tests/test_oracle.py), so the unit draws can be recorded and replayed
on the device (`RECORDER`).
"""
import numpy as np

VERSION2_SCALING = False
# When a list, every call appends (zr, zi): the *unit* normal draws, shape size[:-1] + (K,).
RECORDER = None


def spectrum_scale(exponent, samples, fmin=0.0):
    """Per-bin standard deviation `s_scale[K]` and the normaliser `sigma`."""
    f = np.fft.rfftfreq(samples)
    s_scale = f.copy()
    fmin = max(fmin, 1.0 / samples)
    ix = int(np.sum(s_scale < fmin))
    if ix and ix < len(s_scale):
        s_scale[:ix] = s_scale[ix]
    s_scale = s_scale ** (-exponent / 2.0)
    w = s_scale[1:].copy()
    w[-1] *= (1 + (samples % 2)) / 2.0
    sigma = 2.0 * np.sqrt(np.sum(w ** 2)) / samples
    return s_scale, sigma


def synthesize(zr, zi, exponent, samples, fmin=0.0):
    """Unit normals (..., K) -> coloured series (..., samples); the deterministic half."""
    s_scale, sigma = spectrum_scale(exponent, samples, fmin)
    sr = zr * s_scale
    si = zi * s_scale
    if not (samples % 2):
        si[..., -1] = 0
    si[..., 0] = 0
    if VERSION2_SCALING:
        if not (samples % 2):
            sr[..., -1] *= np.sqrt(2)
        sr[..., 0] *= np.sqrt(2)
    return np.fft.irfft(sr + 1j * si, n=samples, axis=-1) / sigma


def powerlaw_psd_gaussian(exponent, size, fmin=0):
    try:
        size = list(size)
    except TypeError:
        size = [size]
    samples = size[-1]
    size[-1] = samples // 2 + 1
    zr = np.random.standard_normal(size)
    zi = np.random.standard_normal(size)
    if RECORDER is not None:
        RECORDER.append((zr.copy(), zi.copy()))
    return synthesize(zr, zi, exponent, samples, fmin)
