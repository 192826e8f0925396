"""Measured FP64 bars on this GPU (NOT linked into the product): cuBLAS through torch for the shape of the
Schur-complement SYRK (W~^T W~, 1.8 M x 576 at config 4) and a square DGEMM."""
import time, torch
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
W = torch.randn(1800000, 576, dtype=torch.float64, device='cuda')
ms = t(lambda: W.T @ W)
print('cuBLAS dgemm  W^T W  (576 x 1.8M x 576, full product 2*576^2*1.8M = %.0f GFLOP): %.3f ms -> %.2f TFLOP/s' % (2 * 576 ** 2 * 1.8e6 / 1e9, ms, 2 * 576 ** 2 * 1.8e6 / ms / 1e9))
A = torch.randn(8192, 8192, dtype=torch.float64, device='cuda')
ms = t(lambda: A @ A, 3)
print('cuBLAS dgemm 8192^3: %.3f ms -> %.2f TFLOP/s' % (ms, 2 * 8192 ** 3 / ms / 1e9))
