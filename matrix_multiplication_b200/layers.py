'''
``cusparseLinear`` — the FC-layer caller of the hot path (reference
benchmarks/cusparse_fc_layer.py:12-52, which calls ``cusparseMM.apply(inp, self.weight.t())`` and
re-builds a CSR on every step).  Here the *weight* is the sparse operand (BASELINE configs 2-3):
it is held as a persistent device CSR (int32 pattern buffers + an fp32 ``values`` parameter), the
transposed pattern used by the backward pass is built once, and every step runs

    y^T = W_csr @ x^T (+ bias)          forward   : SpMM
    dx^T = W_csr^T @ dy^T               backward  : SpMM over the cached transposed CSR
    dvalues = SDDMM(dy^T, x^T) on W's pattern
'''
import math

import torch
from torch import nn

from . import custom_mm
from .matmuls import CsrMatrix


class _SparseLinearT(torch.autograd.Function):
    '''out[M,N] = W_csr[M,K] @ xT[K,N] for a CsrMatrix whose values are the first argument.'''

    @staticmethod
    def forward(ctx, values, xT, csr):
        csr.refresh_values_if_changed(values)
        ctx.csr = csr
        ctx.save_for_backward(values, xT)
        out = torch.empty(csr.shape[0], xT.shape[1], device=xT.device, dtype=torch.float32)
        return custom_mm.cusparse_mmul(*csr.properties(), xT, out)

    @staticmethod
    def backward(ctx, g):
        values, xT = ctx.saved_tensors
        csr = ctx.csr
        g = g.contiguous()
        d_values = d_xT = None
        if ctx.needs_input_grad[0]:
            d_values = custom_mm.sddmm(csr.crow, csr.col, csr.shape[0], csr.shape[1], g, xT)
        if ctx.needs_input_grad[1]:
            t = csr.transposed()
            d_xT = torch.empty(csr.shape[1], g.shape[1], device=g.device, dtype=torch.float32)
            custom_mm.cusparse_mmul(*t.properties(), g, d_xT)
        return d_values, d_xT, None


class cusparseLinear(nn.Module):
    '''Linear layer ``y = x W^T + b`` with an unstructured-sparse weight kept in CSR.

    :param sparsity: fraction of weight entries that are zero (pattern fixed at construction)
    :param weight: optional dense ``[out_features, in_features]`` tensor whose non-zeros define the layer
    '''

    def __init__(self, in_features, out_features, bias=True, sparsity=0.9, seed=None, weight=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        if weight is None:
            g = torch.Generator().manual_seed(0 if seed is None else seed)
            w = torch.empty(out_features, in_features)
            bound = 1.0 / math.sqrt(in_features)
            w.uniform_(-bound, bound, generator=g)
            w[w == 0] = bound
            w = w * (torch.rand(out_features, in_features, generator=g) >= sparsity)
        else:
            w = weight.detach().to(torch.float32).cpu()
        csr = w.to_sparse_csr()
        self.register_buffer('crow', csr.crow_indices().to(torch.int32))
        self.register_buffer('col', csr.col_indices().to(torch.int32))
        self.values = nn.Parameter(csr.values().clone())
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_features))
        else:
            self.register_parameter('bias', None)
        self._csr = None

    def csr(self) -> CsrMatrix:
        v = self.values
        if self._csr is None or self._csr.crow.data_ptr() != self.crow.data_ptr():
            self._csr = CsrMatrix(self.crow, self.col, v.detach(), (self.out_features, self.in_features))
            self._csr._val_key = (v.data_ptr(), v._version)
        return self._csr

    def forward_t(self, xT):
        '''Transposed-activation form: ``xT`` is [in_features, tokens]; returns [out_features, tokens].
        Chains of sparse layers can stay in this layout and skip the transposes of :meth:`forward`.'''
        out = _SparseLinearT.apply(self.values, xT.contiguous(), self.csr())
        if self.bias is not None:
            out = out + self.bias[:, None]
        return out

    def forward(self, x):
        lead = x.shape[:-1]
        xT = x.reshape(-1, self.in_features).t().contiguous()
        return self.forward_t(xT).t().reshape(*lead, self.out_features)

    def dense_weight(self):
        return torch.sparse_csr_tensor(self.crow, self.col, self.values.detach(),
                                       size=(self.out_features, self.in_features)).to_dense()

    def dense_weight_grad(self):
        return torch.sparse_csr_tensor(self.crow, self.col, self.values.grad,
                                       size=(self.out_features, self.in_features)).to_dense()

    def extra_repr(self):
        return (f'in_features={self.in_features}, out_features={self.out_features}, '
                f'nnz={self.values.numel()}, bias={self.bias is not None}')
