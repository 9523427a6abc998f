"""GraphBatch: the topology of one batch of graphs in the packed HBM layout.

Plays the role of the reference's ``data_slice`` / ``lap_slice`` placeholders
(models/tf_modules/graph_topology.py:64-78,100-135): it records the real node count of every graph
and owns the device-side offset tables (``agcn_plan`` of include/agcn_sgcll.h).
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _stream_ptr(device=None):
    """Current stream OF `device` (not of whatever device happens to be current)."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


class GraphBatch(object):
    def __init__(self, n_nodes, max_atom, device=None):
        n = np.ascontiguousarray(np.asarray(n_nodes, dtype=np.int32).reshape(-1))
        if n.size == 0 or n.min() < 1 or n.max() > max_atom:
            raise ValueError("n_nodes must be in [1, max_atom]")
        self.n_nodes = n
        self.batch_size = int(n.size)
        self.max_atom = int(max_atom)
        self.device = torch.device(device if device is not None else "cuda")
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().agcn_plan_create(n.ctypes.data_as(ctypes.c_void_p), self.batch_size, self.max_atom,
                                                   _stream_ptr(self.device), ctypes.byref(self._handle)))
        self.total_nodes = int(_lib.lib().agcn_plan_total_nodes(self._handle))
        self.total_lap = int(_lib.lib().agcn_plan_total_lap(self._handle))
        self.node_off = np.concatenate([[0], np.cumsum(n, dtype=np.int64)])
        self.lap_off = np.concatenate([[0], np.cumsum(n.astype(np.int64) ** 2)])
        self._graph_ids = None
        self._n_dev = None
        self._host_reads = []   # (event, pinned host tensor) of zero-copy pack kernels still in flight

    @property
    def handle(self):
        return self._handle

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().agcn_plan_destroy(self._handle)
                self._handle = ctypes.c_void_p()
        except Exception:
            pass

    # ---- layout conversion (graph_topology.py:84-98 <-> packed) -------------------------------
    def _device_readable(self, padded):
        """The pack kernels read their input with plain loads, and only the real rows of it: a PINNED host tensor
        (page-locked memory is mapped into the device's address space) can be passed as it is -- the kernel then pulls
        exactly the n_g rows of every graph across PCIe and the zero padding of the wire layout never moves."""
        padded = padded.contiguous().float()
        if not padded.is_cuda and not padded.is_pinned():
            raise ValueError("host tensors must be pinned (tensor.pin_memory()) to be read by the device")
        return padded

    def _hold_host(self, host):
        """A pack kernel is reading `host` (pinned) asynchronously: keep the tensor alive until an event recorded
        behind the kernel has passed, so a temporary (``x.pin_memory()``) cannot be recycled by torch's pinned
        allocator under the kernel.  The CALLER must still not overwrite the buffer before the stream gets there."""
        if host.is_cuda:
            return
        self._host_reads = [(e, t) for e, t in self._host_reads if not e.query()]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._host_reads.append((ev, host))

    def pack_nodes(self, padded):
        """[B, max_atom, F] (device, or pinned host: zero-copy) -> [R, F] (rows >= n_g dropped)."""
        B, N, F = padded.shape
        assert B == self.batch_size and N == self.max_atom
        padded = self._device_readable(padded)
        out = torch.empty(self.total_nodes, F, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().agcn_pack_nodes(self._handle, _ptr(padded), _ptr(out), F, _stream_ptr(self.device)))
            self._hold_host(padded)
        return out

    def unpack_nodes(self, packed):
        """[R, F] -> [B, max_atom, F]; rows >= n_g are exact +0.0 (graphconv.py:249-251)."""
        R, F = packed.shape
        assert R == self.total_nodes
        packed = packed.contiguous()
        out = torch.empty(self.batch_size, self.max_atom, F, device=packed.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().agcn_unpack_nodes(self._handle, _ptr(packed), _ptr(out), F, _stream_ptr(self.device)))
        return out

    def pack_lap(self, padded):
        """[B, max_atom, max_atom] (device, or pinned host: zero-copy) -> packed [sum n^2]."""
        B, N, N2 = padded.shape
        assert B == self.batch_size and N == self.max_atom and N2 == N
        padded = self._device_readable(padded)
        out = torch.empty(self.total_lap, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().agcn_pack_lap(self._handle, _ptr(padded), _ptr(out), _stream_ptr(self.device)))
            self._hold_host(padded)
        return out

    def pack_lap_csr(self, indptr, indices, values):
        """Batch CSR (host numpy: indptr [R+1] over the packed rows, column indices inside each graph, fp32 values)
        -> packed [sum n^2] on the device; the arrays travel through pinned memory."""
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float32)
        if indptr.size != self.total_nodes + 1 or indices.size != values.size or int(indptr[-1]) != indices.size:
            raise ValueError("CSR arrays do not match the batch (indptr needs total_nodes + 1 entries)")
        dev = [torch.from_numpy(a).pin_memory().to(self.device, non_blocking=True) for a in (indptr, indices, values)]
        out = torch.empty(self.total_lap, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().agcn_pack_lap_csr(self._handle, _ptr(dev[0]), _ptr(dev[1]), _ptr(dev[2]), _ptr(out),
                                                    _stream_ptr(self.device)))
        return out

    def point_laplacians(self, points, rule="mean_distance", sparse_ratio=0.1):
        """Packed intrinsic Laplacians of point-cloud graphs built on the device (agcn_point_laplacian): the threshold
        adjacency of meshloader.py:264-285 ("mean_distance") or pointcloudloader.py:240-263 ("cutoff"), then
        Graph.compute_laplacian.  points: packed [R, F] coordinates (device), F <= 8."""
        points = points.contiguous().float()
        R, F = points.shape
        assert R == self.total_nodes and points.is_cuda
        nbytes = ctypes.c_size_t()
        _lib.check(_lib.lib().agcn_point_laplacian_workspace_bytes(self._handle, ctypes.byref(nbytes)))
        work = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        out = torch.empty(self.total_lap, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().agcn_point_laplacian(self._handle, _ptr(points), F, _lib.ADJ_RULE[rule],
                                                       float(sparse_ratio), _ptr(out), _ptr(work), work.numel(),
                                                       _stream_ptr(self.device)))
        return out

    def unpack_lap(self, packed):
        packed = packed.contiguous()
        out = torch.empty(self.batch_size, self.max_atom, self.max_atom, device=packed.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().agcn_unpack_lap(self._handle, _ptr(packed), _ptr(out), _stream_ptr(self.device)))
        return out

    def lap_view(self, packed, g):
        """Unpadded [n_g, n_g] view of graph g (what the reference returns in res_L / res_W / L lists)."""
        n = int(self.n_nodes[g])
        return packed[self.lap_off[g]:self.lap_off[g + 1]].view(n, n)

    def node_view(self, packed, g):
        return packed[self.node_off[g]:self.node_off[g + 1]]

    def n_nodes_device(self):
        """int64 [B] on the device, uploaded from pinned memory (no host synchronisation: a plan per batch
        must not stall a pipelined training loop)."""
        if self._n_dev is None:
            host = torch.from_numpy(self.n_nodes.astype(np.int64)).pin_memory()
            self._n_dev = host.to(self.device, non_blocking=True)
            self._n_host_pinned = host      # keep the staging buffer alive until the copy has run
        return self._n_dev

    def graph_ids(self):
        """int64 [R]: graph index of every packed row (for gathers in the callers)."""
        if self._graph_ids is None:
            self._graph_ids = torch.repeat_interleave(torch.arange(self.batch_size, device=self.device),
                                                      self.n_nodes_device(), output_size=self.total_nodes)
        return self._graph_ids


class PackedNodes(object):
    """List-like stand-in for the reference's ``list of B [max_atom, F] tensors``
    (graphconv.py:118-125): item g is materialised on demand as the zero-padded matrix, while the
    packed storage is what flows between layers."""

    def __init__(self, data, batch):
        self.data = data          # [R, F] torch tensor (autograd-tracked)
        self.batch = batch
        self._padded = None

    def __len__(self):
        return self.batch.batch_size

    def padded(self):
        if self._padded is None:
            from .functional import unpack_nodes
            self._padded = unpack_nodes(self.data, self.batch)
        return self._padded

    def __getitem__(self, g):
        return self.padded()[g]

    def __iter__(self):
        return iter(self.padded().unbind(0))


class PackedLaplacians(object):
    """Lazy list of B unpadded [n_g, n_g] matrices over packed storage (the reference
    fetches every res_L / res_W to the host each step; here nothing moves unless it is read)."""

    def __init__(self, data, batch):
        self.data = data          # [sum n^2]
        self.batch = batch

    def __len__(self):
        return self.batch.batch_size

    def __getitem__(self, g):
        if isinstance(g, slice):
            return [self[i] for i in range(*g.indices(len(self)))]
        if g < 0:
            g += len(self)
        return self.batch.lap_view(self.data, g)

    def __iter__(self):
        return (self[g] for g in range(len(self)))

    def __bool__(self):
        return len(self) > 0
