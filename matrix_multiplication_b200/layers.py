'''
``cusparseLinear`` — the FC-layer caller of the hot path (reference
benchmarks/cusparse_fc_layer.py:12-52, which calls ``cusparseMM.apply(inp, self.weight.t()) + bias`` and
re-builds a CSR on every step).  Here the *weight* is the sparse operand (BASELINE configs 2-3):
it is held as a persistent device CSR (int32 pattern buffers + an fp32 ``values`` parameter) behind an inspected
plan, and every step runs

    y  = act(x W^T + b)                 forward   : ONE engine call — activations read as they are ([tokens, in] is
                                                    B^T), result written as [tokens, out] (C^T), bias and ReLU fused
                                                    into the epilogue: no transposed copies, no separate bias pass
    dx = dy W                           backward  : SpMM over the plan's cached A^T, operands again taken / written
                                                    in [tokens, features] form
    dvalues = SDDMM(dy^T, x^T) on W's pattern

When the weight is too sparse for the block-dense tensor-core tiles (or the batch too small) the layer falls back
to explicit transposes around the gather kernels (bias and ReLU still fused there).
'''
import math

import torch
from torch import nn

from . import custom_mm
from .matmuls import CsrMatrix

_BT, _CT, _RELU = custom_mm.OP_B_TRANSPOSED, custom_mm.OP_C_TRANSPOSED, custom_mm.OP_RELU
_SMALL = 0  # nnz * tokens below which operands are passed as they are to the plain strided kernels: off — measured
            # slower than transposes + gather kernels even on the MNIST MLP (0.33 vs 0.08 ms per graph-replayed step)


class _SparseLinear(torch.autograd.Function):
    '''y[tokens, out] = act(x[tokens, in] @ W^T + bias) for a CsrMatrix W whose values are the first argument.
    Forward and backward are one extension call each (custom_mm.Plan.fc_forward / fc_backward): path choice,
    allocations, ReLU mask and bias gradient all happen on the C++ side.'''

    @staticmethod
    def forward(ctx, values, x, bias, csr, relu):
        csr.refresh_values_if_changed(values)
        small = csr.nnz * x.shape[0] <= _SMALL
        y = csr.plan().fc_forward(x, bias, relu, small)
        ctx.csr, ctx.relu, ctx.has_bias, ctx.small = csr, relu, bias is not None, small
        ctx.save_for_backward(x, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        d_values, d_x, d_bias = ctx.csr.plan().fc_backward(
            g, x, y, ctx.relu, ctx.small, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
            ctx.has_bias and ctx.needs_input_grad[2])
        return d_values, d_x, d_bias, None, None


class cusparseLinear(nn.Module):
    '''Linear layer ``y = act(x W^T + b)`` with an unstructured-sparse weight kept in CSR.

    :param sparsity: fraction of weight entries that are zero (pattern fixed at construction)
    :param weight: optional dense ``[out_features, in_features]`` tensor whose non-zeros define the layer
    :param activation: ``None`` or ``'relu'`` (fused into the product's epilogue)

    State dict: ``crow`` / ``col`` (int32 pattern buffers), ``values``, ``bias``.  A dense ``weight`` entry as saved by
    the reference's layer (benchmarks/cusparse_fc_layer.py:17) is accepted by :meth:`load_state_dict` when its
    non-zero pattern equals this layer's pattern (the pattern is otherwise fixed at construction).
    '''

    def __init__(self, in_features, out_features, bias=True, sparsity=0.9, seed=None, weight=None, activation=None):
        super().__init__()
        assert activation in (None, 'relu')
        self.in_features, self.out_features = in_features, out_features
        self.activation = activation
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
        c = self._csr
        if c is None or c.crow.data_ptr() != self.crow.data_ptr() or c.col.data_ptr() != self.col.data_ptr():
            c = self._csr = CsrMatrix(self.crow, self.col, v.detach(), (self.out_features, self.in_features),
                                      persistent=True)
            c._val_key = (v.data_ptr(), v._version)
        return c

    def values_updated(self):
        '''Call after changing ``values`` through ``.data`` (EMA swaps, manual pruning): such writes do not bump the
        version counter the layer watches.'''
        if self._csr is not None:
            self._csr._val_key = None

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        wkey = prefix + 'weight'
        if wkey in state_dict and prefix + 'values' not in state_dict:  # the reference layer's dense weight
            w = state_dict.pop(wkey).detach().to(torch.float32)
            rows = torch.repeat_interleave(torch.arange(self.out_features), self.crow.cpu().diff().long())
            mask = torch.zeros(self.out_features, self.in_features, dtype=torch.bool)
            mask[rows, self.col.cpu().long()] = True
            if tuple(w.shape) != (self.out_features, self.in_features) or bool(((w.cpu() != 0) & ~mask).any()):
                error_msgs.append(f'{wkey}: non-zero pattern differs from this layer\'s CSR pattern')
            else:
                state_dict[prefix + 'values'] = w.cpu()[rows, self.col.cpu().long()]
                state_dict.setdefault(prefix + 'crow', self.crow)
                state_dict.setdefault(prefix + 'col', self.col)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        self._csr = None  # pattern buffers may have been overwritten in place: drop the plan and its cached transpose

    def forward(self, x):
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.in_features).to(torch.float32)
        if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < x2.shape[1]):
            x2 = x2.contiguous()
        y = _SparseLinear.apply(self.values, x2, self.bias, self.csr(), self.activation == 'relu')
        return y.reshape(*lead, self.out_features)

    def forward_t(self, xT):
        '''Transposed-activation form: ``xT`` is [in_features, tokens]; returns [out_features, tokens].'''
        return self.forward(xT.t()).t()

    def dense_weight(self):
        return torch.sparse_csr_tensor(self.crow, self.col, self.values.detach(),
                                       size=(self.out_features, self.in_features)).to_dense()

    def dense_weight_grad(self):
        return torch.sparse_csr_tensor(self.crow, self.col, self.values.grad,
                                       size=(self.out_features, self.in_features)).to_dense()

    def extra_repr(self):
        return (f'in_features={self.in_features}, out_features={self.out_features}, '
                f'nnz={self.values.numel()}, bias={self.bias is not None}, activation={self.activation}')
