"""
Device-memory plumbing shared by the boundary modules: a process-wide :class:`~nemo_b200.ops.Ops` per precision,
and :class:`DeviceArray`, the lazily-downloaded array the filter classes return in place of numpy arrays so that
``photometry.findObjects`` / ``measureFluxes`` can keep working on the GPU copy while legacy callers that do
``np.array(filteredMapDict['SNMap'])`` still get what they expect.
"""

import os
import numpy as np

_OPS = {}

# bytes moved between host and device through the boundary modules (bench.py reads these per step)
TRAFFIC = {'h2d': 0, 'd2h': 0}


def toHost(t):
    """Device tensor -> numpy array (synchronising copy), counted in TRAFFIC['d2h']."""
    if t.device.type != 'cpu':
        TRAFFIC['d2h'] += int(t.numel()) * t.element_size()
        t = t.detach().cpu()
    return t.numpy()


class HostCopy(object):
    """Device tensor -> pinned host buffer, asynchronous on the current stream; ``.get()`` waits for THIS copy only (an event),
    so kernels queued after it keep the device busy while the host works on the values."""

    def __init__(self, t):
        import torch
        TRAFFIC['d2h'] += int(t.numel()) * t.element_size()
        self.buf = torch.empty(t.shape, dtype=t.dtype, device='cpu', pin_memory=True)
        self.buf.copy_(t.detach(), non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def get(self):
        self.event.synchronize()
        return self.buf.numpy()


def countH2D(t):
    TRAFFIC['h2d'] += int(t.numel()) * t.element_size()
    return t


def defaultPrecision():
    return os.environ.get('NEMO_B200_PRECISION', 'f64')


def getOps(precision=None):
    from .ops import Ops
    if precision is None:
        precision = defaultPrecision()
    import torch
    key = (precision, torch.cuda.current_device() if torch.cuda.is_available() else -1)
    if key not in _OPS:
        _OPS[key] = Ops(precision)
    return _OPS[key]


class DeviceArray(object):
    """A 2-D (or N-D) map living on the GPU with a numpy face.  ``.tensor`` is the torch handle; ``np.asarray``
    (or any numpy function) triggers one device-to-host copy, which is then cached."""

    __array_priority__ = 100

    def __init__(self, tensor, dtype=None):
        self.tensor = tensor
        self._host = None
        self._dtype = dtype

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def ndim(self):
        return self.tensor.dim()

    @property
    def dtype(self):
        if self._dtype is not None:
            return np.dtype(self._dtype)
        return np.dtype(str(self.tensor.dtype).replace('torch.', ''))

    def numpy(self):
        if self._host is None:
            h = toHost(self.tensor)
            if self._dtype is not None:
                h = h.astype(self._dtype)
            self._host = h
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        if dtype is not None:
            a = a.astype(dtype)
        return a

    def __getitem__(self, key):
        return self.numpy()[key]

    def __setitem__(self, key, value):
        # host edit: write through to the device copy so both stay the same map
        a = self.numpy()
        a[key] = value
        import torch
        self.tensor.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(self.tensor.dtype))

    def __len__(self):
        return self.shape[0]

    def sum(self, *a, **k):
        return self.numpy().sum(*a, **k)

    def max(self, *a, **k):
        return self.numpy().max(*a, **k)

    def min(self, *a, **k):
        return self.numpy().min(*a, **k)

    def __mul__(self, other):
        return self.numpy() * np.asarray(other)

    __rmul__ = __mul__

    def __repr__(self):
        return "DeviceArray(shape=%s, dtype=%s, device=%s)" % (self.shape, self.dtype, self.tensor.device)


def asTensor(x, ops, dtype=None):
    """numpy array / DeviceArray / tensor -> contiguous device tensor of the compute dtype (or `dtype`)."""
    if isinstance(x, DeviceArray):
        x = x.tensor
    return ops.to_device(x, dtype)


def gaussFilter(data, sigmaY, sigmaX, subtract=False, ops=None):
    if ops is None:
        ops = getOps()
    wasHost = not (isinstance(data, DeviceArray) or hasattr(data, 'data_ptr'))
    t = asTensor(data, ops)
    squeeze = t.dim() == 2
    if squeeze:
        t = t.unsqueeze(0)
    out = ops.gauss_filter(t.contiguous(), sigmaY, sigmaX, subtract=subtract)
    if squeeze:
        out = out[0]
    if wasHost:
        return toHost(out)
    return out
