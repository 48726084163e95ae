"""CPU check of the folded GELU polynomial used by the GEGLU epilogue (csrc/tc_gemm.cu::gelu_erf) against erf in float64."""
import math
import numpy as np

p = 0.3275911
a = [0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429]
B = 0.5 * (math.sqrt(2.0) / p) * np.poly1d([-1.0, 1.0]) * np.poly1d(a[::-1])
print("B(t) coefficients, highest degree first:", [float(c) for c in B.coeffs])
xs = np.linspace(-8, 8, 200001).astype(np.float32)
t = np.float32(1.0) / (np.abs(xs) * np.float32(p * 0.7071067811865476) + np.float32(1.0))
b = np.float32(B.coeffs[0])
for c in B.coeffs[1:]:
    b = (b * t + np.float32(c)).astype(np.float32)
e = np.exp2((xs * xs) * np.float32(-0.7213475204444817)).astype(np.float32)
got = (np.maximum(xs, 0) - b * e).astype(np.float64)
want = np.array([0.5 * float(x) * (1 + math.erf(float(x) / math.sqrt(2))) for x in xs])
print("max abs err", np.abs(got - want).max(), "rel L2", np.linalg.norm(got - want) / np.linalg.norm(want))
