"""Independent large-N oracle for the GPU tests: a chunked fp32 brute force in plain torch.

TEST INFRASTRUCTURE ONLY (never imported by the product).  The numpy oracle cannot hold BASELINE's
full-size matrices (SURVEY.md hard part e), and checking the library's paths against each other is
not independent.  This class restates ``scores = Qn @ Vn.T ; top-k`` (picovdb/pico_vdb.py:58-68,
584-591, 683-714) with torch ops only -- ``F.normalize``, fp32 ``matmul`` (TF32 off), ``topk`` -- over
the raw rows as they are generated, chunk by chunk, so no full matrix ever exists.  It shares no
code with the CUDA library.  ``tests/test_gpu_parity.py::test_torch_oracle_is_pinned_to_the_numpy_oracle``
pins it to the numpy oracle (itself pinned to the reference's golden vectors) at oracle-sized inputs;
``tests/test_gpu_full_size.py`` then uses it as the expected value at BASELINE's sizes.
"""
from __future__ import annotations

import numpy as np
import torch


class TorchOracle:
    def __init__(self, queries: np.ndarray, k: int, device) -> None:
        assert not torch.backends.cuda.matmul.allow_tf32, "the oracle must multiply in fp32"
        q = torch.as_tensor(np.ascontiguousarray(queries, dtype=np.float32), device=device)
        norms = torch.linalg.vector_norm(q, dim=1, keepdim=True)
        zero = norms[:, 0] == 0                                  # zero query -> e0 (pico_vdb.py:585-590)
        q = torch.where(zero[:, None], torch.zeros_like(q), q)
        q[zero, 0] = 1.0
        norms = torch.where(zero[:, None], torch.ones_like(norms), norms)
        self.qn = (q / norms).contiguous()
        self.k = int(k)
        nq = self.qn.shape[0]
        self.best_s = torch.full((nq, self.k), float("-inf"), device=device)
        self.best_r = torch.full((nq, self.k), -1, dtype=torch.int64, device=device)

    @staticmethod
    def normalize_rows(x: torch.Tensor) -> torch.Tensor:
        n = torch.linalg.vector_norm(x, dim=1, keepdim=True)
        zero = n[:, 0] == 0
        out = x / torch.where(zero[:, None], torch.ones_like(n), n)
        if bool(zero.any()):
            out[zero] = 0.0
            out[zero, 0] = 1.0                                   # zero row -> e0 (pico_vdb.py:58-68)
        return out

    def update(self, raw_rows: torch.Tensor, row0: int, eligible: torch.Tensor | None = None) -> None:
        """Score rows [row0, row0 + len) given RAW (un-normalised) fp32 values; ``eligible`` is an
        optional bool mask over those rows (active & prefilter)."""
        vn = self.normalize_rows(raw_rows.float())
        sc = self.qn @ vn.T
        if eligible is not None:
            sc = sc.masked_fill(~eligible[None, :], float("-inf"))
        kk = min(self.k, sc.shape[1])
        s, r = torch.topk(sc, kk, dim=1)
        r = torch.where(torch.isfinite(s), r + row0, torch.full_like(r, -1))
        cs = torch.cat([self.best_s, s], dim=1)
        cr = torch.cat([self.best_r, r], dim=1)
        # (score descending, row ascending): sort by row first, then stably by score
        order = torch.argsort(torch.where(cr >= 0, cr, torch.full_like(cr, 2**62)), dim=1, stable=True)
        cs, cr = torch.gather(cs, 1, order), torch.gather(cr, 1, order)
        order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, : self.k]
        self.best_s, self.best_r = torch.gather(cs, 1, order), torch.gather(cr, 1, order)

    def result(self) -> tuple[np.ndarray, np.ndarray]:
        return self.best_s.cpu().numpy(), self.best_r.cpu().numpy()
