"""Runs one library op at the bench shapes a few times (for ncu captures and quick timings).
    python tools/run_op.py cell_rms [--precision f64] [--tiles 16] [--bins 8] [--reps 3]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nemo_b200 import device, _lib

ap = argparse.ArgumentParser()
ap.add_argument('op')
ap.add_argument('--precision', default='f64')
ap.add_argument('--tiles', type=int, default=16)
ap.add_argument('--bins', type=int, default=8)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--size', type=int, default=1024)
args = ap.parse_args()
ops = device.getOps(args.precision)
H = W = args.size
nT = args.tiles
g = torch.Generator(device='cuda').manual_seed(1)
maps = torch.randn((nT, H, W), generator=g, device='cuda', dtype=ops.dtype) * 100
yy, xx = torch.meshgrid(torch.arange(H, device='cuda', dtype=ops.dtype), torch.arange(W, device='cuda', dtype=ops.dtype), indexing='ij')
wmed = ((0.75 + 0.25 * torch.sin(2 * np.pi * 1.5 * yy / H + 0.3) * torch.cos(2 * np.pi * xx / W + 1.1)) / 1e4).expand(nT, H, W).contiguous()
grid = 80
yc = np.linspace(0, H, int(H / grid + 1), dtype=int)
xc = np.linspace(0, W, int(W / grid + 1), dtype=int)


psMask = torch.ones((nT, H, W), dtype=torch.uint8, device='cuda')
norm = torch.ones((nT,), dtype=torch.float64, device='cuda')
state = {}


def run():
    if args.op == 'cell_rms':
        return ops.cell_rms(maps, wmed, 1, yc, xc, grid // 2, args.bins)
    if args.op == 'rfft2':
        return ops.rfft2(maps, apodPix=20)
    if args.op == 'apply':
        if 'fm' not in state:
            state['fm'] = ops.rfft2(maps, apodPix=20).reshape(nT, 1, H, W // 2 + 1)
            state['filt'] = torch.rand((nT, 1, H, W // 2 + 1), device='cuda', dtype=ops.dtype)
        return ops.apply_filter(state['fm'], state['filt'], 1, norm=norm, mask=psMask)
    if args.op == 'window':
        return ops.apply_window(maps.clone(), -1.0, restoreZeros=True)
    if args.op == 'snr_mask':
        rms = torch.ones_like(maps)
        return ops.snr_mask(maps.clone(), rms, psMask, psMask, 1, 240, 20, False)
    if args.op == 'filter_build':
        if 'fm' not in state:
            state['fm'] = ops.rfft2(maps, apodPix=20).reshape(nT, 1, H, W // 2 + 1)
            state['fsig'] = ops.rfft2(maps[:1], apodPix=0).reshape(1, 1, H, W // 2 + 1)
        if os.environ.get('TWO_STEP'):
            cov = ops.xpower_smooth(state['fm'], 3.0)
            return ops.build_filter(cov, state['fsig'], [1.0], 1, False)
        return ops.build_filter_fused(state['fm'], state['fsig'], [1.0], 1, False)
    if args.op == 'detect':
        return ops.detect(maps / 100.0, 3.0)
    raise SystemExit("unknown op " + args.op)


run()
torch.cuda.synchronize()
_lib.profEnable(True)
for _ in range(args.reps):
    run()
rep = _lib.profReport()
for k, (n, ms) in rep.items():
    print("%-28s launches %d  avg %.4f ms" % (k, n, ms / n))
