"""GraphTopologyMol: turns a batch of Graph objects into the SGC-LL layer inputs.

Mirrors models/tf_modules/graph_topology.py: ``pad_data2sparse`` (:84-90), ``pad_Lap2sparse``
(:92-98) and ``batch_to_feed_dict`` (:100-135).  Instead of a TensorFlow feed_dict keyed by
placeholders it returns the layer-input dict itself (``node_features``, ``original_laplacian``,
``data_slice``, ``lap_slice``), staged through pinned host memory onto the GPU, plus the batch plan
under ``_batch`` so the layers do not rebuild it.
"""
import numpy as np
import torch

from .batch import GraphBatch, PackedLaplacians, PackedNodes


def batch_csr(laplacians):
    """Per-graph scipy sparse Laplacians (Graph.Laplacian, graph_structure.py:100-107) -> one CSR over the packed rows
    of the batch: ``indptr`` [R + 1] int32, ``indices`` int32 (column inside the row's own graph), ``values`` fp32.
    Duplicate entries are summed first, so every (row, column) appears once (agcn_pack_lap_csr scatters, not adds)."""
    mats = []
    for m in laplacians:
        m = m.tocsr().copy()
        m.sum_duplicates()
        mats.append(m)
    nnz = np.concatenate([[0], np.cumsum([m.nnz for m in mats])]).astype(np.int64)
    if nnz[-1] >= 2 ** 31:
        raise ValueError("batch CSR has more than 2^31 - 1 stored entries")
    indptr = np.concatenate([m.indptr[:-1].astype(np.int64) + o for m, o in zip(mats, nnz[:-1])] + [nnz[-1:]])
    indices = np.concatenate([m.indices for m in mats]).astype(np.int32)
    values = np.concatenate([m.data for m in mats]).astype(np.float32)
    return indptr.astype(np.int32), indices, values


class GraphTopologyMol(object):
    def __init__(self, n_feat, batch_size=50, max_atom=128, name='topology_mol', max_deg=10, min_deg=0,
                 device='cuda'):
        self.n_feat = n_feat
        self.name = name
        self.max_deg = max_deg
        self.min_deg = min_deg
        self.max_atom = max_atom
        self.batch_size = batch_size
        self.device = torch.device(device)

    def pad_data2sparse(self, graph):
        """graph_topology.py:84-90: zero-pad node features to [max_atom, F]; slice = [n, -1]."""
        feature, shape = graph.node_features, graph.node_features.shape
        feature_pad = np.pad(feature, pad_width=((0, self.max_atom - shape[0]), (0, 0)), mode='constant')
        return feature_pad, np.asarray([shape[0], -1], dtype=np.int32)

    def pad_Lap2sparse(self, graph):
        """graph_topology.py:92-98: dense zero-padded Laplacian [max_atom, max_atom]; slice = [n, n]."""
        Laplacian, L_shape = graph.Laplacian, graph.Laplacian.shape
        pad_shape = ((0, self.max_atom - L_shape[0]), (0, self.max_atom - L_shape[1]))
        L_pad = np.pad(np.asarray(Laplacian.todense()), pad_width=pad_shape, mode='constant')
        return L_pad, np.asarray(Laplacian.shape, dtype=np.int32)

    def batch_to_feed_dict(self, batch, layout='packed'):
        """graph_topology.py:100-135.  layout='padded' reproduces the reference's wire format (lists
        of B padded tensors); layout='packed' (default) builds the native HBM layout directly from
        the ragged host arrays, so the padding never crosses PCIe; layout='csr' ships the Laplacians as the CSR
        arrays Graph.compute_laplacian produced and expands them on the device (agcn_pack_lap_csr)."""
        graphs = list(batch)
        n_nodes = np.asarray([g.node_features.shape[0] for g in graphs], dtype=np.int32)
        mol_slice = np.stack([np.asarray([n, -1], dtype=np.int32) for n in n_nodes])
        L_slice = np.stack([np.asarray([n, n], dtype=np.int32) for n in n_nodes])
        plan = GraphBatch(n_nodes, self.max_atom, device=self.device)
        if layout == 'padded':
            feats = np.stack([self.pad_data2sparse(g)[0] for g in graphs]).astype(np.float32)
            laps = np.stack([self.pad_Lap2sparse(g)[0] for g in graphs]).astype(np.float32)
            feats = torch.from_numpy(feats).pin_memory().to(self.device, non_blocking=True)
            laps = torch.from_numpy(laps).pin_memory().to(self.device, non_blocking=True)
            node_features, laplacians = list(feats.unbind(0)), list(laps.unbind(0))
        elif layout == 'packed':
            feats = np.concatenate([np.asarray(g.node_features, np.float32) for g in graphs], 0)
            laps = np.concatenate([np.asarray(g.Laplacian.todense(), np.float32).reshape(-1) for g in graphs])
            feats = torch.from_numpy(feats).pin_memory().to(self.device, non_blocking=True)
            laps = torch.from_numpy(laps).pin_memory().to(self.device, non_blocking=True)
            node_features, laplacians = PackedNodes(feats, plan), PackedLaplacians(laps, plan)
        elif layout == 'csr':
            # Graph.Laplacian is already scipy CSR (graph_structure.py:107): ship its three arrays, expand on the device
            feats = np.concatenate([np.asarray(g.node_features, np.float32) for g in graphs], 0)
            indptr, indices, values = batch_csr([g.Laplacian for g in graphs])
            feats = torch.from_numpy(feats).pin_memory().to(self.device, non_blocking=True)
            laps = plan.pack_lap_csr(indptr, indices, values)
            node_features, laplacians = PackedNodes(feats, plan), PackedLaplacians(laps, plan)
        else:
            raise ValueError("layout must be 'packed', 'csr' or 'padded'")
        return {'node_features': node_features, 'original_laplacian': laplacians,
                'data_slice': mol_slice, 'lap_slice': L_slice, '_batch': plan}
