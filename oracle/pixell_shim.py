"""
Restatement of the four pixell ``enmap`` calls the hot path uses.  pixell (>= 0.17 per the reference's
INSTALL.rst; no lock file) is absent from /root/reference and from this image, so these follow its
published semantics at the reference's call sites.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Call sites restated (all in /root/reference/nemo):
  enmap.fft            filters.py:528,551,562,618,649,682,849
  enmap.ifft           filters.py:665,851          (normalize=False)
  enmap.apod           filters.py:528,551,562,749,849,1154
  enmap.apply_window   filters.py:103,647 ; pipelines.py:231
"""

import numpy as np


def fft(m):
    """``enmap.fft(m)`` with pixell's default ``normalize=True``: fft2 over the last two axes divided
    by sqrt(Npix)."""
    m = np.asarray(m)
    return np.fft.fft2(m, axes=(-2, -1)) / np.sqrt(m.shape[-2] * m.shape[-1])


def ifft(fm, normalize=True):
    """``enmap.ifft``.  ``normalize=False`` (filters.py:665,851) is the plain un-normalised inverse
    sum, i.e. numpy's ifft2 times Npix; ``normalize=True`` is the inverse of :func:`fft`."""
    fm = np.asarray(fm)
    npix = fm.shape[-2] * fm.shape[-1]
    out = np.fft.ifft2(fm, axes=(-2, -1))
    if normalize:
        return out * np.sqrt(npix)
    return out * npix


def apod_profile(width):
    """The cosine taper pixell applies to the first ``width`` samples (reversed for the last ones):
    first sample 0, last sample exactly 1."""
    return 0.5 * (1 - np.cos(np.linspace(0, np.pi, width)))


def apod(m, width):
    """``enmap.apod(m, width)`` (profile='cos', fill='zero')."""
    m = np.array(m, dtype=np.float64, copy=True)
    wy = int(min(width, m.shape[-2]))
    wx = int(min(width, m.shape[-1]))
    if wy > 0:
        a = apod_profile(wy)
        m[..., :wy, :] *= a[:, None]
        m[..., -wy:, :] *= a[::-1, None]
    if wx > 0:
        a = apod_profile(wx)
        m[..., :, :wx] *= a[None, :]
        m[..., :, -wx:] *= a[None, ::-1]
    return m


def calc_window(shape):
    """Order-0 (top-hat pixel) window along each axis: sinc of the FFT frequencies in cycles/pixel."""
    wy = np.sinc(np.fft.fftfreq(shape[-2]))
    wx = np.sinc(np.fft.fftfreq(shape[-1]))
    return wy, wx


def apply_window(m, pow=1.0):
    """``enmap.apply_window(m, pow)``: multiply the 2-D spectrum by the separable pixel window raised
    to ``pow`` and transform back."""
    wy, wx = calc_window(np.shape(m))
    fm = np.fft.fft2(np.asarray(m, dtype=np.float64), axes=(-2, -1))
    fm = fm * (wy[:, None] ** pow)
    fm = fm * (wx[None, :] ** pow)
    return np.fft.ifft2(fm, axes=(-2, -1)).real
