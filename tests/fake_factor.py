"""CPU test double for `B200Factor` (TEST INFRASTRUCTURE ONLY).

Implements the factor's Python surface on top of the oracle so the host-side
drivers in circuitscape_b200/core.py can be exercised on a box with no GPU
(`-m "not gpu"`).  It is never importable from the product package."""
import numpy as np
import scipy.sparse as sp

from oracle import circuitscape_oracle as co


class FakeFactor:
    def __init__(self, matrix, solver, log_transform=False):
        self.A = sp.csr_matrix(matrix, dtype=np.float64)
        self.n = self.A.shape[0]
        self.solver = solver
        self.log = log_transform
        self.dtype = np.dtype(np.float64)
        self.reset_currents()

    def set_grounds(self, finite=None, dirichlet=None):
        """CPU double of cs_b200_set_grounds: diag += finite, identity rows at the Dirichlet nodes."""
        if not hasattr(self, "A0"):
            self.A0 = self.A.copy()
        A = self.A0.tolil(copy=True)
        if finite is not None:
            A.setdiag(A.diagonal() + np.asarray(finite, dtype=np.float64))
        if dirichlet is not None:
            m = np.nonzero(np.asarray(dirichlet))[0]
            A = A.tocsr()
            keep = sp.diags((~np.asarray(dirichlet, dtype=bool)).astype(np.float64))
            A = (keep @ A @ keep + sp.diags(np.asarray(dirichlet, dtype=np.float64))).tolil()
        self.A = sp.csr_matrix(A)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def close(self):
        pass

    def stats(self):
        return {}

    def reset_currents(self):
        self.cum = np.zeros(self.n)
        self.mx = np.full(self.n, -9999.0)

    def read_currents(self, want_max=True):
        return self.cum.copy(), self.mx.copy()

    def solve_pairs(self, src, dst, weight=None, want_volt=False, want_curr=False,
                    accumulate=False, **kw):
        V = co.solve_pairs_direct(self.A, np.asarray(src), np.asarray(dst))
        k = len(src)
        w = np.ones(k) if weight is None else np.asarray(weight, dtype=float)
        curr = np.zeros((self.n, k))
        if accumulate or want_curr:
            for c in range(k):
                cur = co.get_node_currents(self.A, V[:, c])
                curr[:, c] = cur
                if accumulate:
                    val = np.where(cur > 0, np.log10(np.where(cur > 0, cur, 1.0)), -9999.0) if self.log else cur
                    self.cum += w[c] * val
                    self.mx = np.maximum(self.mx, val)
        R = V[np.asarray(dst), np.arange(k)] - V[np.asarray(src), np.arange(k)]
        return dict(R=R, volt=V if want_volt else None, curr=curr if want_curr else None,
                    iters=np.zeros(k, dtype=np.int64), relres=np.zeros(k))

    def solve_pairs_superposed(self, nodes, pi, pj, weight=None, **kw):
        nodes = np.asarray(nodes)
        return self.solve_pairs(nodes[np.asarray(pi)], nodes[np.asarray(pj)], weight, **kw)

    def solve_sources(self, columns, ref, probe=None, weight=None, want_volt=False, want_curr=False,
                      accumulate=False, **kw):
        import scipy.sparse.linalg as spla
        k = len(columns)
        V = np.zeros((self.n, k))
        for c, (rows, vals) in enumerate(columns):
            b = np.zeros(self.n)
            np.add.at(b, np.asarray(rows, dtype=np.int64), np.asarray(vals, dtype=np.float64))
            keep = np.setdiff1d(np.arange(self.n), [ref[c]])
            V[keep, c] = spla.splu(self.A[keep][:, keep].tocsc()).solve(b[keep])
        curr = np.zeros((self.n, k))
        w = np.ones(k) if weight is None else np.asarray(weight, dtype=float)
        if want_curr or accumulate:
            for c in range(k):
                curr[:, c] = co.get_node_currents(self.A, V[:, c])
                if accumulate:
                    self.cum += w[c] * curr[:, c]
                    self.mx = np.maximum(self.mx, curr[:, c])
        pv = None if probe is None else V[np.asarray(probe)].T.copy()
        return dict(probe_volt=pv, volt=V if want_volt else None, curr=curr if want_curr else None,
                    iters=np.zeros(k, dtype=np.int64), relres=np.zeros(k))

    def solve_rhs(self, rhs, **kw):
        import scipy.sparse.linalg as spla
        rhs = np.asarray(rhs, dtype=np.float64)
        x = spla.splu(self.A.tocsc()).solve(rhs)
        k = 1 if rhs.ndim == 1 else rhs.shape[1]
        return x, np.zeros(k, dtype=np.int64), np.zeros(k)
