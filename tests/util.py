"""Helpers shared by the parity tests: match GPU rows to oracle rows by geometric key."""
import numpy as np


def encode_keys(keys):
    k = keys.astype(np.int64)
    return (((k[:, 0] * 3 + k[:, 1]) * 4096 + k[:, 4]) * 4096 + k[:, 3]) * 4096 + k[:, 2]


def perm_gpu_to_oracle(gpu_keys, oracle_keys):
    """perm[i] = oracle row holding the same face as GPU row i (DOF numbering is not part of the contract)."""
    g, o = encode_keys(gpu_keys), encode_keys(oracle_keys)
    assert np.unique(g).size == g.size and np.unique(o).size == o.size
    order = np.argsort(o)
    pos = np.searchsorted(o[order], g)
    assert (pos < o.size).all() and (o[order][pos] == g).all(), "GPU and oracle disagree on the set of velocity DOFs"
    return order[pos]


def csr_permuted(ptr, col, val, perm_rows_to, n):
    """Return scipy CSR of the GPU matrix renumbered into oracle numbering."""
    import scipy.sparse as sp
    A = sp.csr_matrix((val, col, ptr), shape=(n, n)).tocoo()
    return sp.csr_matrix((A.data, (perm_rows_to[A.row], perm_rows_to[A.col])), shape=(n, n))
