"""Times the 2-D FFT passes alone (K1, K4, K8) on a batch of images with the library's per-kernel event timers.
    python tools/time_fft.py --shape 840 1440 --images 128 --freqs 2 --scales 16"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nemo_b200 import device, _lib

ap = argparse.ArgumentParser()
ap.add_argument('--shape', type=int, nargs=2, default=[840, 1440])
ap.add_argument('--tiles', type=int, default=8)
ap.add_argument('--freqs', type=int, default=2)
ap.add_argument('--scales', type=int, default=16)
ap.add_argument('--precision', default='f64')
ap.add_argument('--reps', type=int, default=5)
args = ap.parse_args()
H, W = args.shape
Wc = W // 2 + 1
ops = device.getOps(args.precision)
nT, F, S = args.tiles, args.freqs, args.scales
E = 8 if args.precision == 'f64' else 4
g = torch.Generator(device='cuda').manual_seed(1)
maps = torch.randn((nT * F, H, W), dtype=ops.dtype, device='cuda', generator=g)
filt = torch.rand((nT * S, F, H, Wc), dtype=ops.dtype, device='cuda', generator=g)
mask = torch.ones((nT, H, W), dtype=torch.uint8, device='cuda')
norm = torch.ones((nT * S,), dtype=torch.float64, device='cuda')
def step():
    fm = ops.rfft2(maps, apodPix=20).reshape(nT, F, H, Wc)
    out = ops.apply_filter(fm, filt, S, norm=norm, mask=mask)
    ops.apply_window(out, -1.0, restoreZeros=True)
    return out
for _ in range(2): step()
torch.cuda.synchronize()
_lib.profEnable(True)
for _ in range(args.reps): step()
torch.cuda.synchronize()
rep = _lib.profReport()
_lib.profEnable(False)
N, Nc = H * W, H * Wc
nOut = nT * S
bytesOf = {'rows_fwd': (nT * F + nOut) * (N * E + Nc * 2 * E),
           'rows_inv': 2 * nOut * (Nc * 2 * E + N * E) + nOut * N + nOut * N * E,
           'cols': nT * F * 2 * Nc * 2 * E + nOut * (F * Nc * E + Nc * 2 * E) + nT * F * Nc * 2 * E + nOut * 2 * Nc * 2 * E}
tot = 0
for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    ms /= args.reps
    tot += ms
    key = 'rows_fwd' if 'rows_fwd' in k or 'rfft_rows' in k else 'rows_inv' if 'rows_inv' in k or 'irfft_rows' in k else 'cols' if 'cols' in k else None
    extra = ''
    if key:
        gbs = bytesOf[key] / ms / 1e6
        extra = '  %7.0f GB/s  %.3f of 6527' % (gbs, gbs / 6527.5)
    print("%-28s %4.0f launches %8.3f ms%s" % (k, n / args.reps, ms, extra))
print("total %.3f ms for %d tiles x %d scales of %dx%d, F=%d, %s" % (tot, nT, S, H, W, F, args.precision))
