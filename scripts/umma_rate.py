"""Tensor-pipe rate probe: cycles per 128(256) x 256 x 16 MMA with cta_group::1 and cta_group::2, 2 CTAs and all SMs busy.
Both operands in shared memory (SS mode); in pair mode each CTA holds half of the B operand.
B200: 128.0 cycles per MMA in every case -- the CTA-pair MMA runs at the full rate with half of B per CTA."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from surs_b200 import _capi

ctx = _capi.Context(torch.device("cuda:0"))
for pair in (0, 1):
    for grid in (2, 148):
        c = _capi.selftest_umma_rate(ctx, pair, grid, 4096)
        print("cta_group::%d grid %3d: %.1f cycles per MMA (min %.1f max %.1f over %d issuers)" % (pair + 1, grid, c.mean(), c.min(), c.max(), len(c)))
