"""numpy interpreter of the reduction schedules built by libkge_b200 (kge_build_schedule).

TEST INFRASTRUCTURE ONLY.  It mirrors torchkge_b200/csrc/reduce.cuh operation by operation
(explicit fp32 rounding after every mul / add; fma emulated through float64, which is exact
for the product of two fp32 values) so CPU tests can prove, without a GPU, that replaying a
schedule reproduces ATen's norm(p=1), norm(p=2) and sum(dim=-1) bit for bit.
"""
import numpy as np

F = np.float32

SC_MODE_MASK = 0x03
SC_MODE_A, SC_MODE_T, SC_MODE_T_FMA = 0, 1, 2
SC_CASC1, SC_FOLD1, SC_P_SET, SC_P_ADD, SC_T_ADD_P, SC_T_ADD_A = 0x04, 0x08, 0x10, 0x20, 0x40, 0x80


def _add(a, b):
    return (a.astype(F) + b.astype(F)).astype(F)


def _mul(a, b):
    return (a.astype(F) * b.astype(F)).astype(F)


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def replay(perm, code, values, l2=False):
    """Reduce ``values`` (shape (n_pairs, dim), fp32: the per-index terms, or for l2=True the
    per-index differences x whose squares are summed) in schedule order.  Returns t (n_pairs,).
    """
    n = values.shape[0]
    a = np.zeros(n, F)
    a1 = np.zeros(n, F)
    p = np.zeros(n, F)
    t = np.zeros(n, F)
    for pos in range(len(perm)):
        v = values[:, perm[pos]].astype(F)
        c = int(code[pos])
        mode = c & SC_MODE_MASK
        if l2:
            if mode == SC_MODE_A:
                a = _add(a, _mul(v, v))
            elif mode == SC_MODE_T:
                t = _add(t, _mul(v, v))
            else:
                t = _fma(v, v, t)
        else:
            if mode == SC_MODE_A:
                a = _add(a, v)
            else:
                t = _add(t, v)
        if c & SC_CASC1:
            a1 = _add(a1, a); a = np.zeros(n, F)
        if c & SC_FOLD1:
            a = _add(a, a1); a1 = np.zeros(n, F)
        if c & SC_P_SET:
            p = a.copy(); a = np.zeros(n, F)
        if c & SC_P_ADD:
            p = _add(p, a); a = np.zeros(n, F)
        if c & SC_T_ADD_P:
            t = _add(t, p)
        if c & SC_T_ADD_A:
            t = _add(t, a); a = np.zeros(n, F)
    return t


def bits(x):
    return np.ascontiguousarray(x, dtype=F).view(np.uint32)
