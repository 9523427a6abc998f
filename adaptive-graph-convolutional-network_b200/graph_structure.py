"""Graph / MolGraph: per-sample node features and the intrinsic normalised Laplacian.

Mirrors models/graph_structure.py of the reference (constructor and attributes :22-55, accessors
:57-73, compute_adj_matrix :75-83, compute_laplacian :85-130, MolGraph :133-146).  Host-side
preprocessing, once per sample; built on numpy/scipy only (the reference routes the adjacency
through networkx, which yields the same undirected 0/1 matrix).
"""
import numpy as np
import scipy.sparse as sp


class Graph(object):
    """Adaptive graph object for each data sample."""

    def __init__(self, node_features, adj_lists, max_deg, min_deg):
        self.node_features = node_features
        self.n_node, self.n_feat = node_features.shape
        self.degree_list = np.array([len(nbrs) for nbrs in adj_lists], dtype=np.int32)
        self.adj_lists = adj_lists
        self.max_deg = max_deg
        self.min_deg = min_deg
        self.has_Lap = True
        if self.n_node > 3:                     # graph_structure.py:51-55
            self.Laplacian = self.compute_laplacian(adj_lists)
        else:
            self.Laplacian = None
            self.has_Lap = False

    def get_num_nodes(self):
        return self.n_node

    def get_num_features(self):
        return self.n_feat

    def get_max_degree(self):
        return self.max_deg

    def get_original_adj_list(self):
        return self.adj_lists

    def get_node_fearture(self):
        return self.node_features

    @staticmethod
    def compute_adj_matrix(adj_lists):
        """Adjacency lists -> symmetric 0/1 CSR matrix (an edge listed in either direction counts
        once; self loops stay on the diagonal), node order = list order (graph_structure.py:75-83)."""
        n = len(adj_lists)
        rows, cols = [], []
        for i, nbrs in enumerate(adj_lists):
            for j in nbrs:
                rows += [i, j]
                cols += [j, i]
        A = sp.csr_matrix((np.ones(len(rows), np.int64), (rows, cols)), shape=(n, n))
        A.data[:] = 1                            # duplicates were summed by the constructor
        A.sum_duplicates()
        A.data[:] = 1
        return A

    def compute_laplacian(self, adj_lists):
        """L = I - D^-1/2 A^ D^-1/2 with A^ = D~^-1/2 (A + I) D~^-1/2 (graph_structure.py:85-130).
        float64 CSR, like the reference (sp.eye promotes the float32 adjacency)."""
        adj = self.compute_adj_matrix(adj_lists).astype(np.float32)          # :127
        adj = sp.coo_matrix(adj + sp.eye(adj.shape[0]))                        # :124, :114
        rowsum = np.array(adj.sum(1))
        with np.errstate(divide='ignore'):
            d_inv_sqrt = np.power(rowsum, -0.5).flatten()
        d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0.
        d_mat = sp.diags(d_inv_sqrt)
        W = adj.dot(d_mat).transpose().dot(d_mat)                              # :119
        d = np.asarray(W.sum(axis=0)).squeeze()                                # :93
        d = d + np.spacing(np.array(0, W.dtype))                               # :100
        d = 1 / np.sqrt(d)
        D = sp.diags(d, 0)
        I = sp.identity(d.size, dtype=W.dtype)
        L = sp.csr_matrix(I - D * W * D)                                       # :104
        assert np.abs(L - L.T).mean() < 1e-9                                   # :106
        return L


class MolGraph(Graph):
    """Molecular graph derived from a SMILES string (graph_structure.py:133-146)."""

    def __init__(self, node_features, adj_lists, max_deg=10, min_deg=0, smiles=None):
        super(MolGraph, self).__init__(node_features, adj_lists, max_deg, min_deg)
        self.smiles = smiles
