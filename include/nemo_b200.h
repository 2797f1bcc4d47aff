/*
 * libnemo_b200.so -- C ABI of the B200-native Nemo filter-and-detect hot path.
 *
 * The reference (borisbolliet/nemo-1) is pure Python and has no FFI; the seam this library sits under is
 * the Python call surface pipelines._filterMapsAndMakeCatalogs uses (nemo/pipelines.py:211-214, 251-262,
 * 273-275).  Each entry point below replaces the numpy/scipy/pixell call(s) cited beside it; the Python
 * boundary (nemo_b200/filters.py, nemo_b200/photometry.py) binds them with ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (torch.Tensor.data_ptr()), row-major, caller-owned, unless marked
 *     "host"; nothing is allocated behind the caller's back except FFT twiddle tables cached per (N, prec);
 *   - prec selects the arithmetic/storage type T of maps and spectra: NB_PREC_F32 (0) or NB_PREC_F64 (1);
 *     covariances, filters-in-full-plane, reductions, tables and spline evaluations are always fp64;
 *   - real maps are [B, H, W]; half spectra are [B, H, Wc] complex(T) with Wc = W/2 + 1 (integer division: odd widths too);
 *   - for the FFT entry points H and W must lie in 64..4096 (NB_ERR_UNSUPPORTED otherwise); sides that are powers of two take
 *     direct passes, sides with prime factors 2, 3, 5, 7 that have a compiled plan mixed-radix passes, any other side -- odd
 *     ones included -- exact Bluestein passes (nb_fft_size_supported: 1 / 2 / 0);
 *   - every call is asynchronous on `stream` and returns 0 on success or a negative NB_ERR_* code, with a
 *     message available from nb_last_error() (thread-local).
 */
#ifndef NEMO_B200_H
#define NEMO_B200_H

#include <cuda_runtime.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_PREC_F32 0
#define NB_PREC_F64 1

const char* nb_last_error(void);
int nb_version(void);
int nb_device_info(int* smCount, int* ccMajor, int* ccMinor);
int nb_fft_size_supported(int n);

/* Per-kernel timing with CUDA events (the reference only prints wall-clock "time since start",
 * nemo/pipelines.py:106-107).  nb_prof_enable(1) clears the records and starts timing every kernel this library
 * launches; nb_prof_report synchronises the device and writes "kernel,launches,total_ms\n" lines (first-launch
 * order) into buf (host, cap bytes), returning the full length. */
long long nb_launch_count(void); /* kernels launched by this library since it was loaded */
/* A caller that replays a captured CUDA graph of this library's launches reports the kernels of each replay here, so
 * that nb_launch_count keeps counting kernels that ran. */
long long nb_launch_count_add(long long n);
int nb_prof_enable(int on);
long long nb_prof_report(char* buf, long long cap);

/* K1: enmap.fft(enmap.apod(map, apodPix))  (nemo/filters.py:528, 551, 849; enmap.fft at :618, 649, 682 with
 * apodPix = 0).  out = scale * rfft2(taper * in); the reference's normalisation is scale = 1/sqrt(H*W). */
int nb_rfft2(int prec, const void* in, void* out, int B, int H, int W, int apodPix, double scale,
             cudaStream_t stream);

/* Un-normalised inverse of nb_rfft2: out = scale * norm[b] * mask[b / maskDiv] * irfft2(spec) * (H*W).
 * spec is overwritten (column pass is in place).  norm (fp64, [B]) and mask (u8, [B / maskDiv, H, W]) may be NULL. */
int nb_irfft2(int prec, void* spec, void* out, int B, int H, int W, double scale, const double* norm,
              const unsigned char* mask, int maskDiv, cudaStream_t stream);

/* K4: MatchedFilter.applyFilter (nemo/filters.py:851-857) followed by "* psMask" (:717):
 *   out[o] = scale * norm[o] * mask[o / S] * Re(ifft2_unnormalised(sum_f fmaps[m(o), f] * filt[o, f]))
 * with m(o) = o / S (maps of nOut / S tiles, each filtered at S scales) or, when perScale != 0, m(o) = o % S (S
 * template spectra shared by all tiles: the normalisation pass, filters.py:653, 686).
 * fmaps [nOut / S or S, F, H, Wc] complex(T); filtH [nOut, F, H, Wc] real T = the filter folded onto the half plane,
 * (filt(k) + filt(-k)) / 2, which is what taking the real part of the full complex inverse transform does;
 * scratch [nOut, H, Wc] complex(T).  norm / mask may be NULL.
 * zeroFlags (may be NULL): [nOut] int32, set to 0 for an image in which out + (1 - mask) is nowhere exactly zero -- the pixels
 * the edge filter of nb_snr_mask* searches for (filters.py:730) -- and non-zero otherwise or when the path taken cannot tell (only the
 * row passes with a compile-time plan report it, as a by-product of their store). */
int nb_apply_filter(int prec, const void* fmaps, const void* filtH, void* scratch, void* out, int nOut, int S,
                    int perScale, int F, int H, int W, double scale, const double* norm, const unsigned char* mask,
                    int* zeroFlags, cudaStream_t stream);

/* K8: enmap.apply_window(data, pow) and the zero restore (nemo/filters.py:101-104; pipelines.py:229-232).
 * wy [H], wx [Wc] (type T) hold the window already raised to pow; data [B, H, W] is transformed in place;
 * pixels that were exactly 0 stay 0 when restoreZeros != 0.  scratch [B, H, Wc] complex(T). */
int nb_window_deconv(int prec, void* data, void* scratch, int B, int H, int W, const void* wy, const void* wx,
                     int restoreZeros, cudaStream_t stream);

/* K2: noiseCov[i][j] = ndimage.gaussian_filter(Re(F_i conj(F_j)), (3, 3))  (nemo/filters.py:568-587), on the
 * full [H, W] plane in unshifted layout with scipy's 'reflect' boundary, truncate = 4 (radius 12).
 * fmaps [B, F, H, Wc] complex(T); cov [B, P, H, W] fp64 with P = F (F + 1) / 2 planes in the order
 * (0,0), (0,1), .., (0,F-1), (1,1), ..; tmp same size as cov. */
int nb_xpower_smooth(int prec, const void* fmaps, double* cov, double* tmp, int B, int F, int H, int W,
                     double sigma, cudaStream_t stream);

/* K3: filt[:, y, x] = inv(noiseCov[:, :, y, x]) @ (w * |fSignal[:, y, x]|)  (nemo/filters.py:624-630); exactly
 * singular matrices leave 0, as the reference's try/except does.
 * cov [nOut / S, P, H, W] fp64; fsig [(sigPerTile ? nOut : S), F, H, Wc] complex(T) template spectra;
 * w host [F]; outputs: filtFull [nOut, F, H, W] fp64 (the reference's self.filt; may be NULL) and
 * filtH [nOut, F, H, Wc] real T (folded, for nb_apply_filter). */
int nb_build_filter(int prec, const double* cov, const void* fsig, const double* w, double* filtFull, void* filtH,
                    int nOut, int S, int sigPerTile, int F, int H, int W, cudaStream_t stream);

/* K2 + K3 in one kernel: the filters of nOut / S tiles at S scales straight from the map spectra, the smoothed covariance
 * staying in shared memory (no cov / tmp arrays).  Arguments as nb_xpower_smooth + nb_build_filter. */
int nb_build_filter_fused(int prec, const void* fmaps, const void* fsig, const double* w, double* filtFull, void* filtH,
                          int nOut, int S, int sigPerTile, int F, int H, int W, double sigma, cudaStream_t stream);

/* Fold a full-plane filter (e.g. one loaded from a cache, nemo/filters.py:784-794) onto the half plane. */
int nb_fold_filter(int prec, const double* filtFull, void* filtH, int n, int H, int W, cudaStream_t stream);

/* K5: normalisation scalars (nemo/filters.py:653-670, 686-687).  For each image b of map [B, H, W]:
 * argmaxIdx[b] = np.argmax (first maximum in raster order), maxVal[b] its value, and, when mode == 1 ('yc'),
 * centreVal[b] = bicubic-spline value at (cy, cx) (RectBivariateSpline(...)(cy, cx)); mode 0 skips it.
 * signalNorm[b] = numer / (mode == 1 ? centreVal[b] : maxVal[b]).  scratch: B * 1024 bytes. */
int nb_template_norm(int prec, const void* map, int B, int H, int W, int mode, double cy, double cx, double numer,
                     long long* argmaxIdx, double* maxVal, double* centreVal, double* signalNorm, void* scratch,
                     cudaStream_t stream);

/* fRelWeights (nemo/filters.py:664-670): value at pixel idx[b] of each frequency plane of
 * Re(ifft_unnormalised(fmaps[m(b), f] * filtH[b, f])); m(b) = b when S == 0, else b % S (S template spectra shared
 * by all tiles); out [B, F] fp64; scratch: B * F * 16 doubles.  total (may be NULL): [B] fp64, the value at idx[b] of the sum of
 * the planes (the filtered template itself, which nb_template_norm has just found); the last frequency's plane is then taken as
 * total minus the others instead of a sum over the Fourier plane. */
int nb_plane_values_at(int prec, const void* fmaps, const void* filtH, const long long* idx, double* out, int B,
                       int S, int F, int H, int W, const double* total, double* scratch, cudaStream_t stream);

/* Maps as they are on disk (FITS BITPIX -32; the reference's MapDict.loadTile hands such arrays to MapFilter, nemo/maps.py:83-180,
 * nemo/filters.py:157-164, and numpy widens them in the first arithmetic operation) uploaded as float32 and widened on the
 * device: out[i] = (double) in[i], n elements.  Exact; halves the host-to-device bytes of a tile. */
int nb_widen_f32(const float* in, double* out, long long n, cudaStream_t stream);

/* Per-pixel median over frequencies of the weight maps (nemo/filters.py:360-363). */
int nb_median_weights(int prec, const void* weights, void* out, int B, int F, int H, int W, cudaStream_t stream);

/* MapFilter.makeNoiseMap, noiseGridArcmin = "smart" (nemo/filters.py:366-407): numBins - 1 whole-map bins between the smallest
 * and the largest median weight (np.linspace edges, strict inequalities on both sides), 3-sigma-clipped standard deviation of the
 * non-zero map values of each bin, painted to every pixel of the bin when > 0.
 * map [nImg, H, W]; wmed [nImg / S, H, W].  Outputs: rms [nImg, H, W] T; table [nImg, numBins - 1] fp64 bin RMS values (NULL: not
 * wanted); edges [nImg / S, numBins] fp64.  scratch: 16 * (nImg / S) + 80 * nImg * max(numBins - 1, 1) + 4 * nImg + 64 bytes. */
int nb_smart_rms(int prec, const void* map, const void* wmed, int nImg, int S, int H, int W, int numBins, void* rms, double* table,
                 double* edges, void* scratch, cudaStream_t stream);

/* completeness.getRMSTab (nemo/completeness.py:981-999): map area per distinct noise level of an RMS map.
 * rms [nImg, H, W] T (the masked RMS maps of nb_snr_mask*); mask [nImg, H, W] u8 (the area mask = surveyMaskOut); rowArea
 * [nImg / S, H] fp64 pixel area of every row (maps.getPixelAreaArcmin2Map / 3600); cand [nImg, K] fp64 the levels to look for,
 * ascending, padded with +inf (the values of the noise tables).  Outputs: area [nImg, K] fp64 = sum of rowArea * mask over the
 * pixels at each level; total [nImg] fp64 = masked area; miss [nImg] i32 = non-zero RMS pixels whose value is not in cand. */
int nb_rms_area_table(int prec, const void* rms, const unsigned char* mask, const double* rowArea, int nImg, int S, int H, int W,
                      const double* cand, int K, double* area, double* total, int* miss, cudaStream_t stream);

/* maps.makeModelImage (nemo/maps.py:1734-1899) with maps.makeDegreesDistanceMap (:2414-2471) and signals.makeBeamModelSignalMap
 * (nemo/signals.py:587-619): out[img] = base[img] (or 0 when base is NULL) + sum over the image's objects, in list order, of
 * amplitude * profile(r), r = sqrt(((y - y0) yPixScale)^2 + ((x - x0) xPixScale)^2) inside the object's box [minX, maxX) x
 * [minY, maxY), the profile interpolated linearly as numpy.interp does (0 beyond its last radius).  obj [nObj, 10] fp64: x0, y0,
 * xPixScale, yPixScale, amplitude, minX, maxX, minY, maxY, profile index; objStart [nImg + 1] i32: objects of image img are
 * obj[objStart[img] .. objStart[img + 1]); profR, profV [nProf, nR] fp64 ascending radii (degrees) and profile values.  The pixel
 * window (and for clusters the beam) are applied afterwards with nb_window_deconv / nb_apply_filter. */
int nb_paint_objects(int prec, const void* base, void* out, int nImg, int H, int W, const double* obj, const int* objStart,
                     const double* profR, const double* profV, int nProf, int nR, cudaStream_t stream);

/* out[i] = a[i] + b[i mod nB], i < n: a model image (or one per tile, broadcast over nothing when nB = n) added to the maps it is
 * injected into (nemo/maps.py:2093 sourceInjectionTest: mapDict['data'] = data + modelMap). */
int nb_add_maps(int prec, const void* a, const void* b, void* out, long long n, long long nB, cudaStream_t stream);

/* catalogs.makeOptimalCatalog / removeDuplicates (nemo/catalogs.py:171-183, 520-565; nemo/pipelines.py:329-331): which rows of a
 * stacked catalog lose to a higher-S/N detection of the same object.  raDeg, decDeg, snr [n] fp64 (device).  For every centre row
 * c in [firstCentre, firstCentre + numCentres): the rows within radiusDeg of c by astCoords.calcAngSepDeg(c, row) (c included); if
 * more than one, all but the highest-S/N one (lowest index on ties) get toRemove = 1.  toRemove [n] u8 must be zeroed by the
 * caller; calls over disjoint centre ranges (ranks of a multi-GPU run) are combined by OR.  scratch: nb_xmatch_scratch_ints(n,
 * radiusDeg) int32 words (the rows binned by declination). */
int nb_xmatch_keep_best(const double* raDeg, const double* decDeg, const double* snr, int n, int firstCentre, int numCentres,
                        double radiusDeg, unsigned char* toRemove, int* scratch, cudaStream_t stream);
long long nb_xmatch_scratch_ints(int n, double radiusDeg);

/* K6: MapFilter.makeNoiseMap, grid branch, default 3-sigma-clipped estimator (nemo/filters.py:409-481).
 * map [nImg, H, W] (already multiplied by psMask); wmed [nImg / S, H, W] median weights;
 * yChunks host [ny + 1], xChunks host [nx + 1] (np.linspace(..., dtype=int)); overlapPix; numBins.
 * Outputs: rms [nImg, H, W] T (NULL: no painting, tables only); table [nImg, ny * nx, numBins] fp64 cell RMS values; edges
 * [nImg, ny * nx, numBins + 1] fp64 weight-bin edges (NaN for skipped cells). */
int nb_cell_rms(int prec, const void* map, const void* wmed, int nImg, int S, int H, int W, const int* yChunks,
                int ny, const int* xChunks, int nx, int overlapPix, int numBins, void* rms, double* table,
                double* edges, cudaStream_t stream);

/* nb_cell_rms with the RMS estimator of noiseParams['RMSEstimator'] (nemo/filters.py:460-479): 0 = 3-sigma clipped standard
 * deviation (the default, what nb_cell_rms does), 1 = 'percentile' = np.percentile(abs(binValues), 68.3), 2 = 'biweight' =
 * astropy.stats.biweight_scale(binValues, c = 9.0, modify_sample_size = True) for bins of at least 10 pixels, else 0.
 * Estimators other than 0 run on the streaming kernel (exact order statistics by radix select over the cell). */
int nb_cell_rms_est(int prec, const void* map, const void* wmed, int nImg, int S, int H, int W, const int* yChunks, int ny,
                    const int* xChunks, int nx, int overlapPix, int numBins, int estimator, void* rms, double* table, double* edges,
                    cudaStream_t stream);

/* K7: S/N map, edge trim and final masks (nemo/filters.py:716-758 Fourier class, :1112-1161 real-space class).
 * In: map [nImg, H, W] (filtered * psMask), rms [nImg, H, W], psMask / surveyMask u8 [nImg / S, H, W].
 * edgeCheck = rank_filter(|map + (1 - psMask)|, 0, size = (trimPix, trimPix)) > 0 (all ones when trimPix = 0);
 * Out (may alias the inputs): data = map * edgeCheck (* edgeCheck*surveyMask*psMask unless realSpace),
 * sn = where(rms > 0, map / rms, map) * maskOut with NaN -> 0, rms *= maskOut,
 * maskOut u8 = edgeCheck * surveyMask * psMask * apodMask (apodMask trims apodPix - 1 pixels per edge).
 * scratch: nImg * H * W bytes.  zeroFlags (may be NULL): [nImg] int32 from nb_apply_filter for this very map; an image whose
 * flag is 0 has no zero for the rank filter to find, so edgeCheck is 1 everywhere and neither pass of the filter runs for it. */
int nb_snr_mask(int prec, void* map, void* rms, const unsigned char* psMask, const unsigned char* surveyMask,
                int nImg, int S, int H, int W, int trimPix, int apodPix, int realSpace, void* sn,
                unsigned char* maskOut, unsigned char* scratch, const int* zeroFlags, cudaStream_t stream);

/* K6 painting fused into K7: as nb_snr_mask, but the RMS of every pixel is taken straight from the cell table of
 * nb_cell_rms (called with rms = NULL, which skips its painting pass) instead of from a painted RMS map; rmsOut receives
 * the painted and masked RMS map, or is NULL when nobody reads it (the reference keeps it only with saveRMSMap,
 * filters.py:760-762; here also the noise tables of completeness.getRMSTab and the filter caches of the injection loop). */
int nb_snr_mask_cells(int prec, void* map, const void* wmed, const double* table, const double* edges, const int* yChunks,
                      int ny, const int* xChunks, int nx, int overlapPix, int numBins, const unsigned char* psMask,
                      const unsigned char* surveyMask, int nImg, int S, int H, int W, int trimPix, int apodPix,
                      int realSpace, void* sn, void* rmsOut, unsigned char* maskOut, unsigned char* scratch,
                      const int* zeroFlags, cudaStream_t stream);

/* S/N from a cached RMS map (nemo/pipelines.py:216-228): sn = where(rms > 0, map / rms, map). */
int nb_snr_from_rms(int prec, const void* map, const void* rms, void* sn, int nImg, int S, int H, int W,
                    cudaStream_t stream);

/* K9: segmentation of (map > threshold) with 4-connectivity, numbered exactly like scipy.ndimage.label
 * (raster order of each component's first pixel)  (nemo/photometry.py:212-214).
 * The test is sign * map > threshold (sign = -1 for invertMap, photometry.py:52-53).
 * labels [B, H, W] int32 (0 = background); nLabels [B] int32. */
int nb_label4(int prec, const void* map, double threshold, double sign, int* labels, int* nLabels, int* scratch,
              int B, int H, int W, cudaStream_t stream);
/* number of int32 words nb_label4 needs in scratch */
long long nb_label4_scratch_ints(int B, int H, int W);

/* K10: ndimage.center_of_mass / ndimage.sum / ndimage.maximum_position per label, including label 0
 * (nemo/photometry.py:215-220).  table [B, maxLabels + 1, 8] fp64 rows:
 * {sum v, sum y*v, sum x*v, count, max v, first argmax (linear index), first pixel (linear index), unused};
 * labels above maxLabels are ignored (compare nLabels to detect overflow).  v = sign * map. */
int nb_label_moments(int prec, const void* map, double sign, const int* labels, double* table, int B, int H, int W,
                     int maxLabels, cudaStream_t stream);

/* findObjects' SExtractor-style ellipses (measureShapes = True, nemo/photometry.py:126-178): per label of a segmentation map the
 * corner of its bounding box and the S/N-weighted moments about that corner.  table [B, maxLabels + 1, 8] fp64 rows:
 * {sum d, sum xs d, sum ys d, sum xs^2 d, sum ys^2 d, sum xs ys d, xMin, yMin} with xs = x - xMin, ys = y - yMin, d = sign * map
 * (xMin = yMin = -1 for labels without pixels).  scratch: 2 * B * (maxLabels + 1) int32 words. */
int nb_label_shapes(int prec, const void* map, double sign, const int* labels, double* table, int* scratch, int B, int H, int W,
                    int maxLabels, cudaStream_t stream);

/* K9 + K10 in one asynchronous call, for sparse foregrounds (the usual case above a detection threshold): one dense pass
 * over the map, then linking / numbering / moments over a compact list of the foreground pixels.  Same outputs as
 * nb_label4 + nb_label_moments (labels [B, H, W], nLabels [B], table [B, maxLabels + 1, 8]); flags bit 0 asks for the
 * first-arg-max columns (ndimage.maximum_position, findCenterOfMass = False).  status [B] int32 is set to 1 for an image
 * whose foreground (> H*W/8 pixels) or component count (> 8192) exceeds the compact structures: its outputs are then
 * invalid and the caller must use nb_label4 / nb_label_moments.  scratch: nb_detect_scratch_ints(B, H, W) int32 words. */
int nb_detect(int prec, const void* map, double threshold, double sign, int* labels, int* nLabels, int* status,
              double* table, int maxLabels, int* scratch, int B, int H, int W, int flags, cudaStream_t stream);
long long nb_detect_scratch_ints(int B, int H, int W);

/* The device side of photometry.findObjects in one asynchronous call (nemo/photometry.py:58-70, 101-105): the map is read
 * once for the detection threshold and (haveRing != 0) the ring threshold; both foregrounds are labelled and reduced as in
 * nb_detect; ringFlags [B, maxLabels + 1] u8 is 1 for a detection-threshold label whose position (centre of mass when
 * posMode = 0, first maximum when 1) lies on a ring-threshold component of more than minRingPix pixels that does not contain
 * its own centre of mass.  Outputs for the detection threshold: nLabels [B], table [B, maxLabels + 1, 8] (row 0, the
 * background, is left empty: findObjects never uses it), status [B]
 * (1 = compact structures overflowed at either threshold: use nb_label4 / nb_label_moments / nb_ring_flags), and the
 * dense segmentation labels [B, H, W] when non-NULL.  flags bit 0: first-arg-max columns.
 * scratch: nb_find_objects_scratch_ints(B, H, W) int32 words, 8-byte aligned. */
int nb_find_objects(int prec, const void* map, double threshold, double ringThreshold, int haveRing, double sign,
                    double minRingPix, int posMode, int* nLabels, int* status, double* table, unsigned char* ringFlags,
                    int* labels, int maxLabels, int* scratch, int B, int H, int W, int flags, cudaStream_t stream);
long long nb_find_objects_scratch_ints(int B, int H, int W);

/* Ring removal of findObjects (nemo/photometry.py:60-70, 104-105).  detTable / detN: moment table and label counts at the
 * detection threshold; ringLabels / ringTable: segmentation and moment table at ringThresholdSigma (both from nb_detect or
 * nb_label4 + nb_label_moments, row widths detCap + 1 and ringCap + 1).  flags [B, detCap + 1] u8 = 1 for a candidate whose
 * position (centre of mass when posMode = 0, first maximum when 1) lies on a ring-threshold component of more than
 * minRingPix pixels that does not contain its own centre of mass. */
int nb_ring_flags(const double* detTable, const int* detN, int detCap, const int* ringLabels, const double* ringTable,
                  int ringCap, int B, int H, int W, double minRingPix, int posMode, unsigned char* flags,
                  cudaStream_t stream);

/* K11: interpolate.RectBivariateSpline(arange(H), arange(W), img, kx=3, ky=3)(y, x) at K points
 * (nemo/photometry.py:76-79, 122, 243-253, 295-333; filters.py:657-660).  imgIdx [K] int32 selects the image,
 * yx [K, 2] fp64, out [K] fp64.  Cardinal cubic B-spline interpolation with a truncated (|n| <= 30) exact
 * prefilter; differs from FITPACK's not-a-knot spline only within ~20 pixels of the image edge
 * (by < 0.268^d relative at distance d). */
int nb_spline_eval(int prec, const void* img, int B, int H, int W, const int* imgIdx, const double* yx, double* out,
                   int K, double sign, cudaStream_t stream);

/* K12: ndimage.gaussian_filter(data, (sigmaY, sigmaX)) (reflect, truncate 4), optionally subtracted from the
 * input: out = subtract ? in - smooth : smooth  (nemo/maps.py:1355-1368, 1433-1458).  tmp [B, H, W] T. */
int nb_gauss_filter(int prec, const void* in, void* tmp, void* out, int B, int H, int W, double sigmaY,
                    double sigmaX, int subtract, cudaStream_t stream);

/* K13: out[b] = scale * norm[b] * mask[b] * sum_f ndimage.convolve(in[b, f], kern[f]) (reflect)
 * (nemo/filters.py:1200-1216, 1113).  in [B, F, H, W] T; kern [F, kh, kw] fp64 (odd kh, kw <= 63);
 * planes [B, F, H, W] T receives the per-frequency convolutions when non-NULL (for fRelWeights, :1204-1211). */
int nb_conv2d(int prec, const void* in, const double* kern, void* out, void* planes, int B, int F, int H, int W,
              int kh, int kw, double scale, const double* norm, const unsigned char* mask, cudaStream_t stream);

/* K13 for a batch of (tile, filter) images with a kernel each (the batched real-space driver; filters.py:1200-1216, 1113):
 *   out[b] = scale * norm[b] * mask[b / maskDiv] * sum_f ndimage.convolve(in[m(b), f], kern[k(b), f])   (reflect)
 * m(b) = b % inMod when inMod > 0 (templates shared by the tiles), else b / inDiv (the inDiv filters of a tile);
 * k(b) = b when kernPerImage != 0 (kern [B, F, kh, kw]), else 0 (kern [F, kh, kw]).  Same arithmetic as nb_conv2d,
 * register-blocked (8 outputs per thread). */
int nb_conv2d_batch(int prec, const void* in, const double* kern, void* out, void* planes, int B, int F, int H, int W,
                    int kh, int kw, double scale, const double* norm, const unsigned char* mask, int inDiv, int inMod,
                    int kernPerImage, int maskDiv, cudaStream_t stream);

/* out[b, kx] = mean over ky of in[b, ky, kx] (fp64): the ky average of a folded filter, whose 1-D inverse transform is row 0
 * of the filter's real-space form -- what MapFilter.makeRealSpaceFilterProfile reads (nemo/filters.py:282-304).
 * in [B, H, Wc] T, out [B, Wc] fp64. */
int nb_mean_over_rows(int prec, const void* in, double* out, int B, int H, int Wc, cudaStream_t stream);

/* Small utilities used by the boundary module. */
int nb_cast(int precIn, const void* in, int precOut, void* out, long long n, cudaStream_t stream);
int nb_mask_to_u8(int precIn, const void* in, unsigned char* out, long long n, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
