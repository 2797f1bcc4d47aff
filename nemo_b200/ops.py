"""
Tensor-level wrappers over the C ABI (include/nemo_b200.h).  A :class:`Ops` object fixes the arithmetic type
(fp64 by default, as the reference computes in fp64 when fed float64 maps) and the device; every method takes
and returns torch CUDA tensors, which are used purely as device-memory handles -- all arithmetic happens in
``libnemo_b200.so``.  Calls are asynchronous on torch's current stream.
"""

import ctypes
import numpy as np

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class Ops(object):

    def __init__(self, precision='f64', device=None):
        self.torch = _lib.require_cuda()
        _lib.load()
        torch = self.torch
        if precision in ('f64', 'float64', 'fp64'):
            self.prec, self.dtype, self.cdtype, self.npdtype = _lib.PREC_F64, torch.float64, torch.complex128, np.float64
        elif precision in ('f32', 'float32', 'fp32'):
            self.prec, self.dtype, self.cdtype, self.npdtype = _lib.PREC_F32, torch.float32, torch.complex64, np.float32
        else:
            raise _lib.NemoB200Error("precision must be 'f64' or 'f32' (given %r)" % (precision,))
        self.precision = 'f64' if self.prec == _lib.PREC_F64 else 'f32'
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.launches = 0  # kernels launched through this object (bookkeeping for bench.py's gpu_launches)

    # -- helpers -------------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, shape, dtype=None):
        return self.torch.empty(shape, dtype=dtype if dtype is not None else self.dtype, device=self.device)

    def zeros(self, shape, dtype=None):
        return self.torch.zeros(shape, dtype=dtype if dtype is not None else self.dtype, device=self.device)

    def to_device(self, arr, dtype=None, pinned=False):
        """Host array -> device tensor of the compute dtype (or `dtype`)."""
        torch = self.torch
        if isinstance(arr, torch.Tensor):
            t = arr
        else:
            t = torch.from_numpy(np.ascontiguousarray(arr))
        want = dtype if dtype is not None else self.dtype
        if t.device.type == 'cpu':
            if t.dtype != want:
                t = t.to(want)
            if pinned:
                t = t.pin_memory()
            from . import device as _device
            _device.countH2D(t)
            return t.to(self.device, non_blocking=True)
        return t.to(want).contiguous()

    @staticmethod
    def _width(Wc, W=None, like=None):
        """Map width that goes with half spectra of Wc columns: given, or that of a real array `like`, else the even one."""
        if W is None and like is not None:
            W = int(like.shape[-1])
        if W is None:
            return (Wc - 1) * 2
        if int(W) // 2 + 1 != Wc:
            raise _lib.NemoB200Error("width %d does not go with half spectra of %d columns" % (W, Wc))
        return int(W)

    def _check(self, t, dtype, name):
        if t.dtype != dtype or not t.is_contiguous() or t.device != self.device:
            raise _lib.NemoB200Error("%s must be a contiguous %s tensor on %s (got %s, contiguous=%s, %s)"
                                     % (name, dtype, self.device, t.dtype, t.is_contiguous(), t.device))

    # -- FFTs ----------------------------------------------------------------------------------------
    def rfft2(self, x, apodPix=0, scale=None):
        """enmap.fft(enmap.apod(x, apodPix)) on the half plane: [B, H, W] -> [B, H, W/2+1] complex."""
        self._check(x, self.dtype, 'x')
        B, H, W = x.shape
        if scale is None:
            scale = 1.0 / np.sqrt(float(H) * float(W))
        out = self.empty((B, H, W // 2 + 1), self.cdtype)
        _lib.call('nb_rfft2', self.prec, _ptr(x), _ptr(out), B, H, W, int(apodPix), float(scale), self._stream())
        self.launches += 2
        return out

    def irfft2(self, spec, W, scale=1.0, norm=None, mask=None, maskDiv=1):
        """Un-normalised inverse; `spec` is destroyed."""
        self._check(spec, self.cdtype, 'spec')
        B, H, Wc = spec.shape
        out = self.empty((B, H, W))
        _lib.call('nb_irfft2', self.prec, _ptr(spec), _ptr(out), B, H, W, float(scale), _ptr(norm), _ptr(mask), int(maskDiv),
                  self._stream())
        self.launches += 2
        return out

    def apply_filter(self, fmaps, filtH, S, scale=1.0, norm=None, mask=None, out=None, scratch=None, perScale=False, W=None, zeroFlags=False):
        """out[o] = scale * norm[o] * mask[o // S] * Re(ifft2(sum_f fmaps[o // S, f] * filt[o, f])); with perScale the
        spectra are those of S templates shared by every tile, fmaps[o % S].  zeroFlags=True: returns (out, flags) with flags [nOut]
        int32 = 0 for the images in which out + (1 - mask) is nowhere exactly zero (see nb_apply_filter; for snr_mask*)."""
        self._check(fmaps, self.cdtype, 'fmaps')
        self._check(filtH, self.dtype, 'filtH')
        nT, F, H, Wc = fmaps.shape
        nOut = filtH.shape[0]
        W = self._width(Wc, W, mask)
        if perScale:
            if nT != S or nOut % S != 0 or tuple(filtH.shape[1:]) != (F, H, Wc):
                raise _lib.NemoB200Error("apply_filter: shape mismatch fmaps %s filtH %s S %d" % (tuple(fmaps.shape), tuple(filtH.shape), S))
        elif nOut != nT * S or tuple(filtH.shape[1:]) != (F, H, Wc):
            raise _lib.NemoB200Error("apply_filter: shape mismatch fmaps %s filtH %s S %d" % (tuple(fmaps.shape), tuple(filtH.shape), S))
        if scratch is None:
            scratch = self.empty((nOut, H, Wc), self.cdtype)
        if out is None:
            out = self.empty((nOut, H, W))
        if norm is not None:
            self._check(norm, self.torch.float64, 'norm')
        if mask is not None:
            self._check(mask, self.torch.uint8, 'mask')
        flags = self.empty((nOut,), self.torch.int32) if zeroFlags else None
        _lib.call('nb_apply_filter', self.prec, _ptr(fmaps), _ptr(filtH), _ptr(scratch), _ptr(out), nOut, int(S), 1 if perScale else 0, F, H, W,
                  float(scale), _ptr(norm), _ptr(mask), _ptr(flags), self._stream())
        self.launches += 2
        return (out, flags) if zeroFlags else out

    def window_tables(self, H, W, power):
        """(wy, wx) of enmap.apply_window raised to `power`, on the device (wx on the half plane); cached."""
        cache = self.__dict__.setdefault('_windowTables', {})
        key = (int(H), int(W), float(power))
        if key not in cache:
            cache[key] = self._make_window_tables(H, W, power)
        return cache[key]

    def _make_window_tables(self, H, W, power):
        wy = np.sinc(np.fft.fftfreq(H)) ** power
        wx = (np.sinc(np.fft.fftfreq(W)) ** power)[:W // 2 + 1].copy()
        if W % 2 == 0:
            wx[W // 2] = np.sinc(0.5) ** power  # fftfreq gives -0.5 at the Nyquist column; sinc is even
        return self.to_device(wy), self.to_device(wx)

    def apply_window(self, data, power, restoreZeros=False, tables=None, scratch=None):
        """In-place enmap.apply_window(data, pow=power) (+ zero restore)."""
        self._check(data, self.dtype, 'data')
        B, H, W = data.shape
        if tables is None:
            tables = self.window_tables(H, W, power)
        wy, wx = tables
        if scratch is None:
            scratch = self.empty((B, H, W // 2 + 1), self.cdtype)
        _lib.call('nb_window_deconv', self.prec, _ptr(data), _ptr(scratch), B, H, W, _ptr(wy), _ptr(wx), 1 if restoreZeros else 0,
                  self._stream())
        self.launches += 3
        return data

    # -- filter construction ---------------------------------------------------------------------------
    def xpower_smooth(self, fmaps, sigma=3.0, W=None):
        self._check(fmaps, self.cdtype, 'fmaps')
        B, F, H, Wc = fmaps.shape
        W = self._width(Wc, W)
        P = F * (F + 1) // 2
        cov = self.empty((B, P, H, W), self.torch.float64)
        tmp = self.empty((B, P, H, W), self.torch.float64)
        _lib.call('nb_xpower_smooth', self.prec, _ptr(fmaps), _ptr(cov), _ptr(tmp), B, F, H, W, float(sigma), self._stream())
        self.launches += 2
        return cov

    def build_filter(self, cov, fsig, w, S, sigPerTile, wantFull=False):
        self._check(cov, self.torch.float64, 'cov')
        self._check(fsig, self.cdtype, 'fsig')
        nT, P, H, W = cov.shape
        F = fsig.shape[1]
        nOut = nT * S
        Wc = W // 2 + 1
        filtH = self.empty((nOut, F, H, Wc))
        filtFull = self.empty((nOut, F, H, W), self.torch.float64) if wantFull else None
        warr = (ctypes.c_double * F)(*[float(v) for v in w])
        _lib.call('nb_build_filter', self.prec, _ptr(cov), _ptr(fsig), warr, _ptr(filtFull), _ptr(filtH), nOut, int(S),
                  1 if sigPerTile else 0, F, H, W, self._stream())
        self.launches += 1
        return filtH, filtFull

    def build_filter_fused(self, fmaps, fsig, w, S, sigPerTile, sigma=3.0, wantFull=False, W=None):
        """xpower_smooth + build_filter in one kernel; the covariance never leaves the chip.  Returns (filtH, filtFull or None)."""
        self._check(fmaps, self.cdtype, 'fmaps')
        self._check(fsig, self.cdtype, 'fsig')
        nT, F, H, Wc = fmaps.shape
        W = self._width(Wc, W)
        nOut = nT * S
        filtH = self.empty((nOut, F, H, Wc))
        filtFull = self.empty((nOut, F, H, W), self.torch.float64) if wantFull else None
        warr = (ctypes.c_double * F)(*[float(v) for v in w])
        _lib.call('nb_build_filter_fused', self.prec, _ptr(fmaps), _ptr(fsig), warr, _ptr(filtFull), _ptr(filtH), nOut, int(S),
                  1 if sigPerTile else 0, F, H, W, float(sigma), self._stream())
        self.launches += 1
        return filtH, filtFull

    def fold_filter(self, filtFull):
        self._check(filtFull, self.torch.float64, 'filtFull')
        shp = filtFull.shape
        H, W = shp[-2], shp[-1]
        n = int(np.prod(shp[:-2]))
        filtH = self.empty(tuple(shp[:-1]) + (W // 2 + 1,))
        _lib.call('nb_fold_filter', self.prec, _ptr(filtFull), _ptr(filtH), n, H, W, self._stream())
        self.launches += 1
        return filtH

    def template_norm(self, maps, mode, cy, cx, numer):
        self._check(maps, self.dtype, 'maps')
        B, H, W = maps.shape
        t = self.torch
        argmax = self.empty((B,), t.int64)
        maxVal = self.empty((B,), t.float64)
        centre = self.zeros((B,), t.float64)
        norm = self.empty((B,), t.float64)
        scratch = self.empty((B * 128,), t.float64)
        _lib.call('nb_template_norm', self.prec, _ptr(maps), B, H, W, int(mode), float(cy), float(cx), float(numer), _ptr(argmax),
                  _ptr(maxVal), _ptr(centre), _ptr(norm), _ptr(scratch), self._stream())
        self.launches += 2
        return norm, argmax, maxVal, centre

    def plane_values_at(self, fmaps, filtH, idx, S=0, W=None, total=None):
        """Per-frequency filtered-template values at pixel idx[b]; S > 0: fmaps holds S template spectra shared by all tiles.
        total [B] fp64: the value of the filtered template (all frequencies) at that pixel, when the caller has it -- the planes
        add up to it, so the last frequency's plane is taken as the remainder instead of a sum over the Fourier plane."""
        self._check(fmaps, self.cdtype, 'fmaps')
        self._check(filtH, self.dtype, 'filtH')
        B, F, H, Wc = filtH.shape
        out = self.empty((B, F), self.torch.float64)
        scratch = self.empty((B * F * 16,), self.torch.float64)
        if total is not None:
            self._check(total, self.torch.float64, 'total')
        _lib.call('nb_plane_values_at', self.prec, _ptr(fmaps), _ptr(filtH), _ptr(idx), _ptr(out), B, int(S), F, H, self._width(Wc, W),
                  _ptr(total), _ptr(scratch), self._stream())
        self.launches += 2
        return out

    # -- noise, S/N -------------------------------------------------------------------------------------
    def widen(self, x):
        """float32 device tensor -> float64 tensor of the same shape (maps uploaded in their on-disk format)."""
        torch = self.torch
        if x.dtype != torch.float32 or not x.is_contiguous():
            raise Exception("widen expects a contiguous float32 tensor")
        out = torch.empty(x.shape, dtype=torch.float64, device=x.device)
        _lib.call('nb_widen_f32', _ptr(x), _ptr(out), x.numel(), self._stream())
        self.launches += 1
        return out

    def median_weights(self, weights):
        self._check(weights, self.dtype, 'weights')
        B, F, H, W = weights.shape
        out = self.empty((B, H, W))
        _lib.call('nb_median_weights', self.prec, _ptr(weights), _ptr(out), B, F, H, W, self._stream())
        self.launches += 1
        return out

    ESTIMATORS = {'clip': 0, 'percentile': 1, 'biweight': 2}

    def cell_rms(self, maps, wmed, S, yChunks, xChunks, overlapPix, numBins, estimator='clip', paint=True):
        """Cell RMS tables (estimator 'clip': 3-sigma clipped standard deviation; 'percentile': 68.3th percentile of |v|;
        'biweight': astropy's biweight_scale with c = 9 and the sample size counted inside the cut, 0 below 10 pixels) and
        (paint=True) the painted RMS map; with paint=False returns (None, table, edges)."""
        self._check(maps, self.dtype, 'maps')
        self._check(wmed, self.dtype, 'wmed')
        nImg, H, W = maps.shape
        ny, nx = len(yChunks) - 1, len(xChunks) - 1
        rms = self.empty((nImg, H, W)) if paint else None
        table = self.empty((nImg, ny * nx, numBins), self.torch.float64)
        edges = self.empty((nImg, ny * nx, numBins + 1), self.torch.float64)
        yc = (ctypes.c_int * (ny + 1))(*[int(v) for v in yChunks])
        xc = (ctypes.c_int * (nx + 1))(*[int(v) for v in xChunks])
        if estimator not in self.ESTIMATORS:
            raise Exception("RMSEstimator '%s' is not implemented in nemo_b200 (clip, percentile, biweight)" % estimator)
        if estimator == 'clip':
            _lib.call('nb_cell_rms', self.prec, _ptr(maps), _ptr(wmed), nImg, int(S), H, W, yc, ny, xc, nx, int(overlapPix), int(numBins),
                      _ptr(rms), _ptr(table), _ptr(edges), self._stream())
        else:
            _lib.call('nb_cell_rms_est', self.prec, _ptr(maps), _ptr(wmed), nImg, int(S), H, W, yc, ny, xc, nx, int(overlapPix),
                      int(numBins), self.ESTIMATORS[estimator], _ptr(rms), _ptr(table), _ptr(edges), self._stream())
        self.launches += 2 if paint else 1
        return rms, table, edges

    def smart_rms(self, maps, wmed, S, numBins):
        """RMS map of the 'smart' noise option: whole-map weight bins.  Returns (rms [nImg, H, W], table [nImg, numBins - 1],
        edges [nImg / S, numBins])."""
        self._check(maps, self.dtype, 'maps')
        self._check(wmed, self.dtype, 'wmed')
        torch = self.torch
        nImg, H, W = maps.shape
        numBins = int(numBins)
        nb = max(numBins - 1, 1)
        rms = self.empty((nImg, H, W))
        table = self.zeros((nImg, nb), torch.float64)
        edges = self.empty((nImg // int(S), numBins), torch.float64)
        scratch = self.empty((16 * (nImg // int(S)) + 80 * nImg * nb + 4 * nImg + 64,), torch.uint8)
        _lib.call('nb_smart_rms', self.prec, _ptr(maps), _ptr(wmed), nImg, int(S), H, W, numBins, _ptr(rms), _ptr(table), _ptr(edges),
                  _ptr(scratch), self._stream())
        self.launches += 28 if numBins > 1 else 4
        return rms, table[:, :numBins - 1], edges

    def rms_area_table(self, rms, mask, rowArea, S, cand):
        """Area per noise level: rms [nImg, H, W], mask u8 [nImg, H, W], rowArea [nImg / S, H] f64, cand [nImg, K] f64 ascending,
        +inf padded.  Returns (area [nImg, K], total [nImg], miss [nImg] i32)."""
        self._check(rms, self.dtype, 'rms')
        self._check(mask, self.torch.uint8, 'mask')
        self._check(rowArea, self.torch.float64, 'rowArea')
        self._check(cand, self.torch.float64, 'cand')
        nImg, H, W = rms.shape
        K = cand.shape[1]
        area = self.empty((nImg, K), self.torch.float64)
        total = self.empty((nImg,), self.torch.float64)
        miss = self.empty((nImg,), self.torch.int32)
        _lib.call('nb_rms_area_table', self.prec, _ptr(rms), _ptr(mask), _ptr(rowArea), nImg, int(S), H, W, _ptr(cand), K, _ptr(area),
                  _ptr(total), _ptr(miss), self._stream())
        self.launches += 1
        return area, total, miss

    def paint_objects(self, shape, objects, objStart, profR, profV, base=None, out=None):
        """Model images [nImg, H, W]: objects [nObj, 10] fp64 (x0, y0, xPixScale, yPixScale, amplitude, minX, maxX, minY, maxY, profile
        index; host array or device tensor), objStart [nImg + 1] i32, profile tables profR / profV [nProf, nR] fp64; added to `base`
        [nImg, H, W] when given."""
        t = self.torch
        nImg, H, W = shape
        obj = objects if hasattr(objects, 'data_ptr') else self.to_device(np.ascontiguousarray(objects, dtype=np.float64).reshape(-1, 10), t.float64)
        start = objStart if hasattr(objStart, 'data_ptr') else self.to_device(np.ascontiguousarray(objStart, dtype=np.int32), t.int32)
        pr = profR if hasattr(profR, 'data_ptr') else self.to_device(np.ascontiguousarray(np.atleast_2d(profR), dtype=np.float64), t.float64)
        pv = profV if hasattr(profV, 'data_ptr') else self.to_device(np.ascontiguousarray(np.atleast_2d(profV), dtype=np.float64), t.float64)
        if tuple(pr.shape) != tuple(pv.shape) or int(start.shape[0]) != nImg + 1:
            raise _lib.NemoB200Error("paint_objects: profile tables %s / %s or objStart %s do not fit" % (tuple(pr.shape), tuple(pv.shape), tuple(start.shape)))
        if obj.shape[0] == 0:
            obj = self.zeros((1, 10), t.float64)
        if base is not None:
            self._check(base, self.dtype, 'base')
        if out is None:
            out = self.empty((nImg, H, W))
        else:
            self._check(out, self.dtype, 'out')
        _lib.call('nb_paint_objects', self.prec, _ptr(base), _ptr(out), int(nImg), int(H), int(W), _ptr(obj), _ptr(start), _ptr(pr), _ptr(pv),
                  int(pr.shape[0]), int(pr.shape[1]), self._stream())
        self.launches += 1
        return out

    def add_maps(self, a, b, out=None):
        """a + b, b broadcast over the leading axes of a (a model image per tile added to every frequency map of the tile...)."""
        self._check(a, self.dtype, 'a')
        self._check(b, self.dtype, 'b')
        if a.numel() % b.numel() != 0:
            raise _lib.NemoB200Error("add_maps: sizes %d and %d do not fit" % (a.numel(), b.numel()))
        if out is None:
            out = self.empty(tuple(a.shape))
        else:
            self._check(out, self.dtype, 'out')
        _lib.call('nb_add_maps', self.prec, _ptr(a), _ptr(b), _ptr(out), int(a.numel()), int(b.numel()), self._stream())
        self.launches += 1
        return out

    def xmatch_keep_best(self, RADeg, decDeg, SNR, radiusDeg, firstCentre=0, numCentres=None):
        """catalogs.makeOptimalCatalog's decision for host columns RADeg, decDeg, SNR [n]: device u8 mask [n], 1 where a row loses
        to a higher-S/N row within radiusDeg of some centre in [firstCentre, firstCentre + numCentres) (default: all rows)."""
        ra = self.to_device(np.ascontiguousarray(RADeg, dtype=np.float64), self.torch.float64)
        dec = self.to_device(np.ascontiguousarray(decDeg, dtype=np.float64), self.torch.float64)
        snr = self.to_device(np.ascontiguousarray(SNR, dtype=np.float64), self.torch.float64)
        n = int(ra.shape[0])
        toRemove = self.zeros((n,), self.torch.uint8)
        scratch = self.empty((max(1, int(_lib.load().nb_xmatch_scratch_ints(n, float(radiusDeg)))),), self.torch.int32)
        _lib.call('nb_xmatch_keep_best', _ptr(ra), _ptr(dec), _ptr(snr), n, int(firstCentre), int(n if numCentres is None else numCentres),
                  float(radiusDeg), _ptr(toRemove), _ptr(scratch), self._stream())
        self.launches += 4
        return toRemove

    def snr_mask(self, maps, rms, psMask, surveyMask, S, trimPix, apodPix, realSpace, zeroFlags=None):
        """In place on maps / rms; returns (data, SNMap, surveyMaskOut u8, rms)."""
        self._check(maps, self.dtype, 'maps')
        self._check(rms, self.dtype, 'rms')
        self._check(psMask, self.torch.uint8, 'psMask')
        self._check(surveyMask, self.torch.uint8, 'surveyMask')
        nImg, H, W = maps.shape
        sn = self.empty((nImg, H, W))
        maskOut = self.empty((nImg, H, W), self.torch.uint8)
        scratch = self.empty((nImg, H, W), self.torch.uint8) if trimPix > 0 else None
        if zeroFlags is not None:
            self._check(zeroFlags, self.torch.int32, 'zeroFlags')
        _lib.call('nb_snr_mask', self.prec, _ptr(maps), _ptr(rms), _ptr(psMask), _ptr(surveyMask), nImg, int(S), H, W, int(trimPix),
                  int(apodPix), 1 if realSpace else 0, _ptr(sn), _ptr(maskOut), _ptr(scratch), _ptr(zeroFlags), self._stream())
        self.launches += 2 if trimPix > 0 else 1
        return maps, sn, maskOut, rms

    def snr_mask_cells(self, maps, wmed, table, edges, yChunks, xChunks, overlapPix, numBins, psMask, surveyMask, S, trimPix, apodPix,
                       realSpace, wantRMS=True, zeroFlags=None):
        """snr_mask with the RMS taken straight from the cell tables (painting fused); returns (data, SNMap, mask u8, RMS map);
        the RMS map is None with wantRMS=False (8 bytes per pixel less to write when nothing reads it)."""
        self._check(maps, self.dtype, 'maps')
        self._check(wmed, self.dtype, 'wmed')
        self._check(psMask, self.torch.uint8, 'psMask')
        self._check(surveyMask, self.torch.uint8, 'surveyMask')
        nImg, H, W = maps.shape
        ny, nx = len(yChunks) - 1, len(xChunks) - 1
        sn = self.empty((nImg, H, W))
        rms = self.empty((nImg, H, W)) if wantRMS else None
        maskOut = self.empty((nImg, H, W), self.torch.uint8)
        scratch = self.empty((nImg, H, W), self.torch.uint8) if trimPix > 0 else None
        yc = (ctypes.c_int * (ny + 1))(*[int(v) for v in yChunks])
        xc = (ctypes.c_int * (nx + 1))(*[int(v) for v in xChunks])
        _lib.call('nb_snr_mask_cells', self.prec, _ptr(maps), _ptr(wmed), _ptr(table), _ptr(edges), yc, ny, xc, nx, int(overlapPix),
                  int(numBins), _ptr(psMask), _ptr(surveyMask), nImg, int(S), H, W, int(trimPix), int(apodPix), 1 if realSpace else 0,
                  _ptr(sn), _ptr(rms), _ptr(maskOut), _ptr(scratch), _ptr(zeroFlags), self._stream())
        self.launches += 2 if trimPix > 0 else 1
        return maps, sn, maskOut, rms

    def snr_from_rms(self, maps, rms, S=1):
        self._check(maps, self.dtype, 'maps')
        self._check(rms, self.dtype, 'rms')
        nImg, H, W = maps.shape
        sn = self.empty((nImg, H, W))
        _lib.call('nb_snr_from_rms', self.prec, _ptr(maps), _ptr(rms), _ptr(sn), nImg, int(S), H, W, self._stream())
        self.launches += 1
        return sn

    # -- detection ---------------------------------------------------------------------------------------
    def label4(self, maps, threshold, sign=1.0):
        self._check(maps, self.dtype, 'maps')
        B, H, W = maps.shape
        t = self.torch
        labels = self.empty((B, H, W), t.int32)
        nLabels = self.empty((B,), t.int32)
        nScratch = _lib.load().nb_label4_scratch_ints(B, H, W)
        scratch = self.empty((int(nScratch),), t.int32)
        _lib.call('nb_label4', self.prec, _ptr(maps), float(threshold), float(sign), _ptr(labels), _ptr(nLabels), _ptr(scratch), B, H, W,
                  self._stream())
        self.launches += 6
        return labels, nLabels

    DETECT_MAX_LABELS = 8192

    def detect(self, maps, threshold, sign=1.0, maxLabels=None, wantArgmax=False):
        """Threshold + 4-connected labelling + per-label moments in one asynchronous call (sparse foreground path).
        Returns device tensors (labels, nLabels, status, table); status[b] != 0 means image b overflowed the compact
        structures and label4 / label_moments must be used instead."""
        self._check(maps, self.dtype, 'maps')
        B, H, W = maps.shape
        t = self.torch
        if maxLabels is None:
            maxLabels = self.DETECT_MAX_LABELS
        labels = self.empty((B, H, W), t.int32)
        nLabels = self.empty((B,), t.int32)
        status = self.empty((B,), t.int32)
        table = self.empty((B, int(maxLabels) + 1, 8), t.float64)
        scratch = self.empty((int(_lib.load().nb_detect_scratch_ints(B, H, W)),), t.int32)
        _lib.call('nb_detect', self.prec, _ptr(maps), float(threshold), float(sign), _ptr(labels), _ptr(nLabels), _ptr(status), _ptr(table),
                  int(maxLabels), _ptr(scratch), B, H, W, 1 if wantArgmax else 0, self._stream())
        self.launches += 8
        return labels, nLabels, status, table

    def find_objects(self, maps, threshold, ringThreshold=None, sign=1.0, minRingPix=30, findCenterOfMass=True, maxLabels=None,
                     wantLabels=False):
        """Device side of findObjects in one asynchronous call: returns (nLabels, status, table, ringFlags or None, labels or
        None) as device tensors; see nb_find_objects."""
        self._check(maps, self.dtype, 'maps')
        B, H, W = maps.shape
        t = self.torch
        if maxLabels is None:
            maxLabels = self.DETECT_MAX_LABELS
        haveRing = ringThreshold is not None
        nLabels = self.empty((B,), t.int32)
        status = self.empty((B,), t.int32)
        table = self.empty((B, int(maxLabels) + 1, 8), t.float64)
        ringFlags = self.empty((B, int(maxLabels) + 1), t.uint8) if haveRing else None
        labels = self.empty((B, H, W), t.int32) if wantLabels else None
        nInts = int(_lib.load().nb_find_objects_scratch_ints(B, H, W))
        scratch = self.empty(((nInts + 1) // 2,), t.int64)       # 8-byte aligned
        _lib.call('nb_find_objects', self.prec, _ptr(maps), float(threshold), float(ringThreshold) if haveRing else 0.0, 1 if haveRing else 0,
                  float(sign), float(minRingPix), 0 if findCenterOfMass else 1, _ptr(nLabels), _ptr(status), _ptr(table), _ptr(ringFlags),
                  _ptr(labels), int(maxLabels), _ptr(scratch), B, H, W, 0 if findCenterOfMass else 1, self._stream())
        self.launches += 14 if haveRing else 7
        return nLabels, status, table, ringFlags, labels

    def ring_flags(self, detTable, detN, ringLabels, ringTable, minRingPix=30, findCenterOfMass=True):
        """u8 [B, detCap + 1]: 1 where a candidate sits on a ring (photometry.py:60-70, 104-105); asynchronous."""
        t = self.torch
        self._check(detTable, t.float64, 'detTable')
        self._check(ringTable, t.float64, 'ringTable')
        self._check(ringLabels, t.int32, 'ringLabels')
        self._check(detN, t.int32, 'detN')
        B, H, W = ringLabels.shape
        detCap, ringCap = detTable.shape[1] - 1, ringTable.shape[1] - 1
        flags = self.empty((B, detCap + 1), t.uint8)
        _lib.call('nb_ring_flags', _ptr(detTable), _ptr(detN), detCap, _ptr(ringLabels), _ptr(ringTable), ringCap, B, H, W,
                  float(minRingPix), 0 if findCenterOfMass else 1, _ptr(flags), self._stream())
        self.launches += 1
        return flags

    def label_moments(self, maps, labels, maxLabels, sign=1.0):
        self._check(maps, self.dtype, 'maps')
        self._check(labels, self.torch.int32, 'labels')
        B, H, W = maps.shape
        table = self.empty((B, int(maxLabels) + 1, 8), self.torch.float64)
        _lib.call('nb_label_moments', self.prec, _ptr(maps), float(sign), _ptr(labels), _ptr(table), B, H, W, int(maxLabels), self._stream())
        self.launches += 4
        return table

    def label_shapes(self, maps, labels, maxLabels, sign=1.0):
        """[B, maxLabels + 1, 8] fp64: bounding-box corner and S/N-weighted moments of every label (nb_label_shapes)."""
        self._check(maps, self.dtype, 'maps')
        self._check(labels, self.torch.int32, 'labels')
        B, H, W = maps.shape
        table = self.empty((B, int(maxLabels) + 1, 8), self.torch.float64)
        scratch = self.empty((2 * B * (int(maxLabels) + 1),), self.torch.int32)
        _lib.call('nb_label_shapes', self.prec, _ptr(maps), float(sign), _ptr(labels), _ptr(table), _ptr(scratch), B, H, W, int(maxLabels),
                  self._stream())
        self.launches += 4
        return table

    def spline_eval(self, imgs, imgIdx, yx, sign=1.0):
        self._check(imgs, self.dtype, 'imgs')
        B, H, W = imgs.shape
        K = int(yx.shape[0])
        out = self.empty((K,), self.torch.float64)
        if K == 0:
            return out
        self._check(yx, self.torch.float64, 'yx')
        if imgIdx is not None:
            self._check(imgIdx, self.torch.int32, 'imgIdx')
        _lib.call('nb_spline_eval', self.prec, _ptr(imgs), B, H, W, _ptr(imgIdx), _ptr(yx), _ptr(out), K, float(sign), self._stream())
        self.launches += 1
        return out

    # -- real-space path -----------------------------------------------------------------------------------
    def gauss_filter(self, maps, sigmaY, sigmaX, subtract=False):
        self._check(maps, self.dtype, 'maps')
        shp = maps.shape
        H, W = shp[-2], shp[-1]
        B = int(np.prod(shp[:-2]))
        tmp = self.empty(shp)
        out = self.empty(shp)
        _lib.call('nb_gauss_filter', self.prec, _ptr(maps), _ptr(tmp), _ptr(out), B, H, W, float(sigmaY), float(sigmaX),
                  1 if subtract else 0, self._stream())
        self.launches += 3 if subtract else 2
        return out

    def conv2d(self, maps, kern, scale=1.0, norm=None, mask=None, wantPlanes=False):
        self._check(maps, self.dtype, 'maps')
        self._check(kern, self.torch.float64, 'kern')
        B, F, H, W = maps.shape
        kh, kw = kern.shape[-2], kern.shape[-1]
        out = self.empty((B, H, W))
        planes = self.empty((B, F, H, W)) if wantPlanes else None
        _lib.call('nb_conv2d', self.prec, _ptr(maps), _ptr(kern), _ptr(out), _ptr(planes), B, F, H, W, kh, kw, float(scale), _ptr(norm),
                  _ptr(mask), self._stream())
        self.launches += 1
        return (out, planes) if wantPlanes else out

    def conv2d_batch(self, maps, kern, nOut, scale=1.0, norm=None, mask=None, inDiv=1, inMod=0, maskDiv=1, wantPlanes=False):
        """K13 for nOut (tile, filter) images: kern [nOut, F, kh, kw] (one kernel per image) or [F, kh, kw]; output image b reads
        maps[b % inMod] (inMod > 0) or maps[b // inDiv], and mask[b // maskDiv]."""
        self._check(maps, self.dtype, 'maps')
        self._check(kern, self.torch.float64, 'kern')
        nIn, F, H, W = maps.shape
        kh, kw = kern.shape[-2], kern.shape[-1]
        perImage = 1 if kern.dim() == 4 else 0
        if perImage and kern.shape[0] != nOut:
            raise Exception("conv2d_batch: %d kernels for %d images" % (kern.shape[0], nOut))
        out = self.empty((nOut, H, W))
        planes = self.empty((nOut, F, H, W)) if wantPlanes else None
        _lib.call('nb_conv2d_batch', self.prec, _ptr(maps), _ptr(kern), _ptr(out), _ptr(planes), int(nOut), F, H, W, kh, kw, float(scale),
                  _ptr(norm), _ptr(mask), int(inDiv), int(inMod), perImage, int(maskDiv), self._stream())
        self.launches += 1
        return (out, planes) if wantPlanes else out

    def mean_over_rows(self, planes):
        """[B, H, Wc] -> [B, Wc] fp64: mean over the rows of every plane."""
        self._check(planes, self.dtype, 'planes')
        B, H, Wc = planes.shape
        out = self.empty((B, Wc), self.torch.float64)
        _lib.call('nb_mean_over_rows', self.prec, _ptr(planes), _ptr(out), B, H, Wc, self._stream())
        self.launches += 1
        return out

    def scale_mask(self, maps, scale=1.0, norm=None, mask=None):
        """maps * scale * norm[b] * mask[b] via a 1x1 convolution (keeps all arithmetic in the library)."""
        # the 1 x 1 kernel is made once per Ops on the device: a host-to-device copy from pageable memory is not allowed while a
        # stream is being captured into a CUDA graph (the batched bckSub path runs under capture)
        one = getattr(self, '_unitKernel', None)
        if one is None:
            one = self._unitKernel = self.torch.ones((1, 1, 1), dtype=self.torch.float64, device=self.device)
        return self.conv2d(maps.unsqueeze(1).contiguous(), one, scale=scale, norm=norm, mask=mask)
