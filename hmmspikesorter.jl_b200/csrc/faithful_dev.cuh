// faithful_dev.cuh -- the reference's scalar arithmetic with explicit rounding (no FMA contraction)
#pragma once
namespace hmm {
// Gaussian log-emission, src/utils.jl:3-4: (-log2pi - l_sigma) - (dd*dd)/(2*sigma2)
__device__ __forceinline__ double emit_rn(double x, double mu, double c_emit, double two_s2) {
    double dd = __dsub_rn(x, mu);
    return __dsub_rn(c_emit, __ddiv_rn(__dmul_rn(dd, dd), two_s2));
}
}  // namespace hmm
