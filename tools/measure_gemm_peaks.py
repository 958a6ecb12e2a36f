#!/usr/bin/env python
"""Yardstick only: what cuBLAS reaches on this GPU for a large square GEMM in TF32 and bf16.

MEASURED_PEAKS.json (driver-written) holds the bf16 figure the rooflines use; it has no TF32 entry,
so the TF32 rows of DESIGN.md are quoted against bf16/2 *and* against the number printed here.
Not part of the product path.
"""
import json

import torch


def gemm_tflops(dtype, n=8192, iters=30, tf32=False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(5):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return 2.0 * n ** 3 / best / 1e9


def gemm_burst_and_sustained(dtype, n=8192, tf32=False, seconds=4.0):
    """The driver's method for MEASURED_PEAKS.json: best single launch of 10 (burst) and the mean
    over a back-to-back loop of `seconds` (sustained, under the power cap)."""
    import time

    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, launches = time.perf_counter(), 0
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        launches += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    flops = 2.0 * n ** 3
    return flops / best / 1e9, flops * launches / e0.elapsed_time(e1) / 1e9


if __name__ == "__main__":
    tf32_burst, tf32_sus = gemm_burst_and_sustained(torch.float32, tf32=True)
    bf16_burst, bf16_sus = gemm_burst_and_sustained(torch.bfloat16)
    out = {
        "tf32_tflops_burst": tf32_burst, "tf32_tflops_sustained": tf32_sus,
        "bf16_tflops_burst": bf16_burst, "bf16_tflops_sustained": bf16_sus,
        "gpu": torch.cuda.get_device_name(0),
        "cublas_tf32_tflops_8192": gemm_tflops(torch.float32, tf32=True),
        "cublas_tf32_tflops_16384x": gemm_tflops(torch.float32, n=12288, iters=10, tf32=True),
        "cublas_bf16_tflops_8192": gemm_tflops(torch.bfloat16),
        "cublas_fp32_simt_tflops_4096": gemm_tflops(torch.float32, n=4096, iters=5, tf32=False),
    }
    print(json.dumps(out))
