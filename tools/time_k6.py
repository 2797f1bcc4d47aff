"""Times K6 (nb_cell_rms) alone on synthetic filtered maps: Gaussian noise with an apodised-looking zero border and a few bright
sources, constant or varying weights.
    python tools/time_k6.py --shape 840 1440 --images 128 --scales 16 --bins 1
NEMO_B200_LIB=build_exp/lib_<tag>.so selects a variant build (tools/build_variant.sh)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nemo_b200 import device, _lib

ap = argparse.ArgumentParser()
ap.add_argument('--shape', type=int, nargs=2, default=[840, 1440])
ap.add_argument('--images', type=int, default=128)
ap.add_argument('--scales', type=int, default=16)
ap.add_argument('--bins', type=int, default=1)
ap.add_argument('--grid', type=int, default=80)
ap.add_argument('--precision', default='f64')
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--sources', type=int, default=40)
args = ap.parse_args()
H, W = args.shape
ops = device.getOps(args.precision)
dt = torch.float64 if args.precision == 'f64' else torch.float32
g = torch.Generator(device='cuda').manual_seed(1)
nT = max(1, args.images // args.scales)
nImg = nT * args.scales
maps = torch.randn((nImg, H, W), generator=g, device='cuda', dtype=dt)
yy, xx = np.mgrid[0:H, 0:W]
rng = np.random.default_rng(2)
src = np.zeros((H, W))
for _ in range(args.sources):
    y0, x0 = rng.integers(0, H), rng.integers(0, W)
    src += rng.uniform(5, 60) * np.exp(-((yy - y0) ** 2 + (xx - x0) ** 2) / (2 * 2.5 ** 2))
maps += torch.from_numpy(src).to('cuda', dt)[None]
maps[:, :20, :] = 0
maps[:, -20:, :] = 0
maps[:, :, :20] = 0
maps[:, :, -20:] = 0
wmed = (1.0 + 0.3 * torch.rand((nT, H, W), generator=g, device='cuda', dtype=dt)).contiguous()
n = lambda L: np.linspace(0, L, int(L / args.grid + 1), dtype=int)
yc, xc, ov = n(H), n(W), args.grid // 2
for _ in range(2):
    ops.cell_rms(maps, wmed, args.scales, yc, xc, ov, args.bins, paint=False)
torch.cuda.synchronize()
_lib.profEnable(True)
for _ in range(args.reps):
    ops.cell_rms(maps, wmed, args.scales, yc, xc, ov, args.bins, paint=False)
rep = _lib.profReport()
_lib.profEnable(False)
peak = 6527.5
byt = nImg * H * W * maps.element_size()
for k, (cnt, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    per = ms / cnt
    print("%-28s %8.3f ms per launch   %7.0f GB/s of map bytes  %.3f of %.0f" % (k, per, byt / per / 1e6, byt / per / 1e6 / peak, peak))
lib = _lib.load()
if hasattr(lib, 'nb_onebin_stats'):
    import ctypes
    out = (ctypes.c_ulonglong * 4)()
    lib.nb_onebin_stats(out, 1)
    c = max(1, out[0])
    print("one-bin stats per cell: sweeping passes %.3f, list passes %.3f, list overflows %.4f (%d cells)" % (out[1] / c, out[2] / c, out[3] / c, out[0]))
print("total %.3f ms for %d images of %dx%d (%d cells each), %d bin(s), %s" % (sum(ms for _, ms in rep.values()) / args.reps, nImg, H, W,
                                                                             (len(yc) - 1) * (len(xc) - 1), args.bins, args.precision))
