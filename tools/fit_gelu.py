"""Fit of the one-MUFU GELU used by the GEGLU epilogue (csrc/gemm_tc.cu: gelu_erf):
   gelu(x) = max(x, 0) - a * 2^Q(a),  a = min(|x|, 5.5),  Q = degree-6 minimax fit of log2 Phi(-a) on [0, 5.5].
Prints the coefficients and checks the fp32 evaluation against the exact function on every finite fp16 input.
   python tools/fit_gelu.py"""
import math

import numpy as np
from scipy.special import log_ndtr, ndtr

A, DEG = 5.5, 6
n = 8000
a = (np.cos(np.pi * (np.arange(n) + 0.5) / n) * 0.5 + 0.5) * A
y = log_ndtr(-a) / math.log(2.0)
V = np.vander(a, DEG + 1, increasing=True)
w = np.ones(n)
for _ in range(300):                      # Lawson iteration towards the minimax fit
    c, *_ = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)
    e = np.abs(V @ c - y)
    w = w * (e / e.mean() + 1e-12) ** 0.5
    w /= w.mean()
print("coefficients (a^0 .. a^6):", [float(np.float32(v)) for v in c])
c32 = c.astype(np.float32)


def gelu_fast(x):
    x = x.astype(np.float32)
    ac = np.minimum(np.abs(x), np.float32(A))
    q = np.full_like(x, c32[DEG])
    for k in range(DEG - 1, -1, -1):
        q = (q * ac + c32[k]).astype(np.float32)
    t = np.exp2(q.astype(np.float64)).astype(np.float32)
    return (np.maximum(x, np.float32(0)) - ac * t).astype(np.float32)


h = np.arange(0, 65536, dtype=np.uint16).view(np.float16)
h = h[np.isfinite(h)].astype(np.float64)
ref = h * ndtr(h)
got = gelu_fast(h).astype(np.float64)
err = np.abs(got - ref)
print("all finite fp16 inputs: max abs error %.3e, max error relative to max(|gelu|, 1e-3): %.3e"
      % (err.max(), (err / np.maximum(np.abs(ref), 1e-3)).max()))
print("results that round to a different fp16 value:", int((ref.astype(np.float16) != got.astype(np.float16)).sum()),
      "of", len(h))
