"""Offline GPTQ conversion of a dense `torch.nn.Linear` into a `ColBlockQuantizedLinear`
(reference: `GPTQQuantizer`, lit_llama/quantization.py:425-616, driven by quantize/gptq.py:84-120;
algorithm: Frantar et al., "GPTQ: Accurate Post-Training Quantization for Generative Pre-trained
Transformers", arXiv:2210.17323).

This is the step *before* the decode path: it produces the checkpoint the
kernels consume.  It is host-side linear algebra on whatever device the layer lives on (torch
library ops: one Cholesky factorisation and rank-1 updates per layer) and is not on the per-token path.

What the algorithm does, per layer with weight W (out, in) and calibration inputs X:
  1. H = 2/n * sum x x^T  (running mean over batches, `collect_input_stats` is a forward hook);
  2. choose the per-row (or per-group) affine grid from W: scale = (max - min) / maxq, zero = round(-min / scale);
  3. walk the input columns in blocks; round column j to the grid, divide the rounding error by the
     j-th diagonal entry of U (U = upper Cholesky factor of H^-1) and push it onto the not-yet-rounded
     columns with row j of U, inside the block immediately and onto the later blocks once per block;
  4. pack the rounded weight with `ColBlockQuantizedLinear.pack_weight`.
Same constructor, attributes and return value as the reference class, and the same order of floating
point operations, so a conversion reproduces the reference's integers (tests/test_gptq_cpu.py).
"""
import math
from typing import Tuple

import torch

from .quantization import ColBlockQuantizedLinear


def affine_grid(block: torch.Tensor, maxq: int, *, per_row: bool = True, symmetric: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """(scale, zero) of the asymmetric min/max grid of every row of `block` (quantization.py:477-513).
    The range always contains 0; an all-zero row gets the range [-1, 1]."""
    rows = block.shape[0]
    flat = block.flatten(1) if per_row else block.flatten().unsqueeze(0)
    origin = torch.zeros(flat.shape[0], device=block.device)
    lo = torch.minimum(flat.min(1)[0], origin)
    hi = torch.maximum(flat.max(1)[0], origin)
    if symmetric:
        hi = torch.maximum(lo.abs(), hi)
        neg = lo < 0
        if torch.any(neg):
            lo[neg] = -hi[neg]
    empty = (lo == 0) & (hi == 0)
    lo[empty] = -1
    hi[empty] = +1
    scale = (hi - lo) / maxq
    zero = torch.full_like(scale, (maxq + 1) / 2) if symmetric else torch.round(-lo / scale)
    if not per_row:
        scale, zero = scale.repeat(rows), zero.repeat(rows)
    tail = [1] * (block.dim() - 1)
    return scale.reshape(-1, *tail), zero.reshape(-1, *tail)


def snap_to_grid(x: torch.Tensor, scale: torch.Tensor, zero: torch.Tensor, maxq: int) -> torch.Tensor:
    """The grid point nearest to x, as a float: scale * (clamp(round(x / scale) + zero, 0, maxq) - zero)."""
    level = torch.clamp(torch.round(x / scale) + zero, 0, maxq)
    return scale * (level - zero)


def error_feedback_factor(H: torch.Tensor, percdamp: float) -> torch.Tensor:
    """Upper Cholesky factor of (H + damp I)^-1 (quantization.py:551-557): row j holds, up to the diagonal
    entry, how the rounding error of column j is distributed over the columns that follow."""
    n = H.shape[0]
    idx = torch.arange(n, device=H.device)
    H[idx, idx] += percdamp * torch.mean(torch.diag(H))
    lower = torch.linalg.cholesky(H)
    return torch.linalg.cholesky(torch.cholesky_inverse(lower), upper=True)


class GPTQQuantizer:
    """Drop-in for lit_llama.quantization.GPTQQuantizer (same signature, attributes and results)."""

    def __init__(self, linear_module, *, bits, perchannel=True, sym=False, blocksize=128, percdamp=0.01, groupsize=-1,
                 actorder=False):
        assert isinstance(linear_module, torch.nn.Linear)
        assert not (actorder and groupsize != -1), "The permutation trick does not work for grouped quantization"
        self.linear_module = linear_module
        w = linear_module.weight
        self.dev = w.device
        self.rows, self.columns = w.shape
        self.H = torch.zeros((self.columns, self.columns), device=self.dev)
        self.nsamples = 0
        self.bits, self.maxq = bits, 2**bits - 1
        self.perchannel, self.sym = perchannel, sym
        self.blocksize, self.percdamp = blocksize, percdamp
        self.groupsize, self.actorder = groupsize, actorder
        self.tile_cols = self.columns if groupsize == -1 else groupsize
        n_groups = (self.columns + self.tile_cols - 1) // self.tile_cols
        self.scales = torch.zeros((self.rows, n_groups), dtype=w.dtype, device=self.dev)
        self.zeros = torch.zeros_like(self.scales)

    # the two helpers the reference exposes under these names (quantize/gptq.py does not call them, tests do)
    @staticmethod
    def quantize_weight(x, scale, zero, maxq):
        return snap_to_grid(x, scale, zero, maxq)

    def find_params_weight(self, x):
        return affine_grid(x, self.maxq, per_row=self.perchannel, symmetric=self.sym)

    def collect_input_stats(self, _module, inp, _out):
        """Forward hook: fold one calibration batch into H = (2 / n) * sum of x x^T (quantization.py:515-529)."""
        x = inp[0].detach()
        self.last_inp = x
        if x.dim() == 2:
            x = x.unsqueeze(0)
        batch = x.shape[0]
        x = x.reshape(-1, x.shape[-1]).t()                 # (features, tokens)
        self.H *= self.nsamples / (self.nsamples + batch)  # running mean: shrink what is there ...
        self.nsamples += batch
        x = math.sqrt(2 / self.nsamples) * x.float()       # ... and add the new batch with weight 2 / n
        self.H += x.matmul(x.t())

    def quantize(self):
        """-> (ColBlockQuantizedLinear, summed weighted squared rounding error)."""
        lin = self.linear_module
        W = lin.weight.detach().to(dtype=torch.float, copy=True)
        scale, zero = self.find_params_weight(W)
        self.scales[:] = scale
        self.zeros[:] = zero

        H = self.H
        del self.H
        unused = torch.diag(H) == 0          # input features the calibration data never excited
        H[unused, unused] = 1
        W[:, unused] = 0
        order = None
        if self.actorder:                    # most "active" input features first
            order = torch.argsort(torch.diag(H), descending=True)
            W = W[:, order]
            H = H[order][:, order]
        U = error_feedback_factor(H, self.percdamp)

        rounded = torch.zeros_like(W)
        loss = torch.zeros_like(W)
        for start in range(0, self.columns, self.blocksize):
            stop = min(start + self.blocksize, self.columns)
            Wb = W[:, start:stop].clone()
            Ub = U[start:stop, start:stop]
            Qb = torch.zeros_like(Wb)
            Eb = torch.zeros_like(Wb)
            Lb = torch.zeros_like(Wb)
            for j in range(stop - start):
                col = start + j
                if self.groupsize != -1 and col % self.groupsize == 0:
                    # a new group starts: its grid comes from the error-compensated weights of the group
                    # (the reference raises here, quantization.py:578 assigns a (rows, 1) tensor to a (rows,) column)
                    scale, zero = self.find_params_weight(W[:, col : col + self.groupsize])
                    self.scales[:, col // self.groupsize] = scale.squeeze(1)
                    self.zeros[:, col // self.groupsize] = zero.squeeze(1)
                w = Wb[:, j]
                pivot = Ub[j, j]
                q = snap_to_grid(w.unsqueeze(1), scale, zero, self.maxq).squeeze(1)
                assert q.dim() == 1
                Qb[:, j] = q
                Lb[:, j] = (w - q) ** 2 / pivot**2
                e = (w - q) / pivot
                Wb[:, j:] -= e.unsqueeze(1).matmul(Ub[j, j:].unsqueeze(0))   # rank-1: spread the error inside the block
                Eb[:, j] = e
            rounded[:, start:stop] = Qb
            loss[:, start:stop] = Lb / 2
            W[:, stop:] -= Eb.matmul(U[start:stop, stop:])                   # ... and onto every later block at once
        if order is not None:
            rounded = rounded[:, torch.argsort(order)]

        weight = rounded.reshape(lin.weight.shape).to(lin.weight.data.dtype)
        error = torch.sum(loss).item()
        q_module = ColBlockQuantizedLinear(lin.in_features, lin.out_features, lin.bias is not None, bits=self.bits,
                                           tile_cols=self.groupsize).to(self.dev)
        q_module.scales = self.scales
        q_module.zeros = self.zeros
        q_module.pack_weight(weight)
        q_module.bias = lin.bias
        return q_module, error
