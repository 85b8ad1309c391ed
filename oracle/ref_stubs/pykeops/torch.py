"""TEST INFRASTRUCTURE ONLY - dense torch stand-in for pykeops.torch.LazyTensor.

Implements exactly the six operations the reference loss uses
(/root/reference/src/losses/focus.py:129-137,159): ``-``, ``** 2``, ``.abs()``,
``.sum(-1)``, ``.argKmin(K, dim)`` and ``.Kmin(K, axis)``, with the KeOps shape
convention: a [.., n, 1, 2] operand is the i-variable, a [.., 1, q, 2] operand
the j-variable, the symbolic result has shape [.., n, q] and a reduction over
``dim=2`` (the i axis of a [B, nb, n, q] tensor) returns [B, nb, q, K].

The KNN is evaluated densely but chunked over the q axis so a 19200x19200 slab
(1.5 GB in f32) is never materialised.  Tie-breaking: the lowest trajectory
index wins on equal distance (stable sort) - KeOps' own rule is not observable
here (parity of the *neighbour set* under exact ties is therefore unpinned).
"""
import torch


class LazyTensor:
    def __init__(self, x=None, *, _a=None, _b=None, _op=None):
        self._x = x
        self._a, self._b, self._op = _a, _b, _op

    # ---- symbolic algebra (only what focus.py needs) -------------------------
    def __sub__(self, other):
        return LazyTensor(_a=self, _b=other, _op='sub')

    def __pow__(self, p):
        assert p == 2
        return LazyTensor(_a=self, _op='sq')

    def abs(self):
        return LazyTensor(_a=self, _op='abs')

    def sum(self, dim):
        assert dim in (-1, 4)
        return LazyTensor(_a=self, _op='sum')

    # ---- evaluation on a q-chunk ------------------------------------------
    def _leaves(self):
        if self._op is None:
            return [self]
        out = self._a._leaves()
        if self._b is not None:
            out += self._b._leaves()
        return out

    def _eval(self, sl):
        if self._op is None:
            x = self._x
            # j-variable: [..., 1, q, 2] -> slice q
            if x.shape[-3] == 1 and x.shape[-2] != 1:
                return x[..., :, sl, :]
            return x
        if self._op == 'sub':
            return self._a._eval(sl) - self._b._eval(sl)
        if self._op == 'sq':
            return self._a._eval(sl) ** 2
        if self._op == 'abs':
            return self._a._eval(sl).abs()
        if self._op == 'sum':
            return self._a._eval(sl).sum(-1)
        raise NotImplementedError(self._op)

    def _dims(self):
        n = q = None
        batch = ()
        for leaf in self._leaves():
            x = leaf._x
            if x.shape[-3] == 1 and x.shape[-2] != 1:
                q = x.shape[-2]                       # j-variable [.., 1, q, 2]
            else:
                n = x.shape[-3]                       # i-variable [.., n, 1, 2]
            if x.dim() - 3 > len(batch):
                batch = tuple(x.shape[:-3])
        return batch, n, q

    @property
    def shape(self):
        batch, n, q = self._dims()
        return tuple(batch) + (n, q)

    def _kmin(self, K, dim, want_index):
        batch, n, q = self._dims()
        assert dim == len(batch), "only a reduction over the i (trajectory) axis is used"
        nbatch = 1
        for b in batch:
            nbatch *= b
        chunk = max(1, min(q, (1 << 25) // max(1, n * nbatch)))
        outs = []
        with torch.no_grad():
            for s in range(0, q, chunk):
                d = self._eval(slice(s, s + chunk))            # [.., n, qc]
                d = d.transpose(-1, -2)                        # [.., qc, n]
                if n <= 4096:
                    vals, idx = torch.sort(d, dim=-1, stable=True)  # lowest index wins ties
                    vals, idx = vals[..., :K], idx[..., :K]
                else:
                    # big problems: top-(K+8), then order (value, index) - same rule unless
                    # more than 8 exact ties straddle the K-th place
                    kk = min(n, K + 8)
                    vals, idx = torch.topk(d, kk, dim=-1, largest=False, sorted=True)
                    key = torch.argsort(idx, dim=-1, stable=True)
                    vals, idx = vals.gather(-1, key), idx.gather(-1, key)
                    key = torch.argsort(vals, dim=-1, stable=True)
                    vals, idx = vals.gather(-1, key)[..., :K], idx.gather(-1, key)[..., :K]
                outs.append(idx if want_index else vals)
        return torch.cat(outs, dim=-2)                         # [.., q, K]

    def argKmin(self, K, dim):
        return self._kmin(K, dim, True)

    def Kmin(self, K, axis):
        return self._kmin(K, axis, False)
