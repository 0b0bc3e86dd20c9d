// mdg_common.cuh — device helpers shared by the sm_100a kernels: error plumbing, Philox4x32-10
// streams, warp/sub-warp reductions and the FP64 special functions (lgamma + digamma in one pass).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "mdg.h"

namespace mdg {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define MDG_CUDA_TRY(expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            mdg::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                \
                           cudaGetErrorString(_e));                                            \
            return MDG_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, keyed by (seed, tax_id); counter layout (c0 draw index, c1 iteration,
// c2 run_kind | purpose << 8, c3 aux) — see DESIGN.md "Random streams"
// ---------------------------------------------------------------------------------------------
enum Purpose : uint32_t { P_INIT = 1, P_MOM = 2, P_DIR = 3, P_SUB = 4, P_HEUR = 5, P_PPC = 6 };

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__host__ __device__ __forceinline__ uint2 make_key(uint64_t seed, int64_t tax_id) {
    uint64_t k = splitmix64(seed ^ splitmix64((uint64_t)tax_id));
    return make_uint2((uint32_t)k, (uint32_t)(k >> 32));
}

__device__ __forceinline__ uint4 philox4x32(uint2 key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = key.x, k1 = key.y;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t c2word(int run_kind, uint32_t purpose) {
    return (uint32_t)run_kind | (purpose << 8);
}

// two uniforms in (0,1): 53 random bits + half an ulp
__device__ __forceinline__ void uniform2(uint4 o, double& u0, double& u1) {
    u0 = ((double)(o.x >> 5) * 67108864.0 + (double)(o.y >> 6) + 0.5) * (1.0 / 9007199254740992.0);
    u1 = ((double)(o.z >> 5) * 67108864.0 + (double)(o.w >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

// two standard normals (Box-Muller)
__device__ __forceinline__ void normal2(uint4 o, double& n0, double& n1) {
    double u0, u1;
    uniform2(o, u0, u1);
    double rad = sqrt(-2.0 * log(u0));
    double s, c;
    sincos(6.283185307179586476925 * u1, &s, &c);
    n0 = rad * c;
    n1 = rad * s;
}

// ---------------------------------------------------------------------------------------------
// sub-warp butterflies: every lane of a GW-wide group ends with the bit-identical sum
// ---------------------------------------------------------------------------------------------
template <int GW>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
    for (int o = GW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

template <int GW>
__device__ __forceinline__ unsigned group_mask() {
    if constexpr (GW == 32) {
        return 0xffffffffu;
    } else {
        unsigned lane = threadIdx.x & 31u;
        return ((1u << GW) - 1u) << (lane & ~(unsigned)(GW - 1));
    }
}

// ---------------------------------------------------------------------------------------------
// FP64 special functions
// ---------------------------------------------------------------------------------------------

// lgamma(x) and digamma(x) for x > 0 in one pass. x < 10 is shifted to y = x + 10 with the
// product P = x(x+1)...(x+9) and its derivative P' (so that lgamma(x) = lgamma(y) - log P and
// digamma(x) = digamma(y) - P'/P: one log and one division instead of ten); y >= 10 uses the
// Stirling / asymptotic series with 7 Bernoulli terms (truncation < 4e-17 absolute at y = 10).
__device__ __forceinline__ void lgam_digam(double x, double& lg, double& dg) {
    double y = x, logP = 0.0, dP = 0.0;
    if (x < 10.0) {
        double P = x, Q = 1.0;
#pragma unroll
        for (int i = 1; i < 10; ++i) {
            double xi = x + (double)i;
            Q = fma(Q, xi, P);
            P *= xi;
        }
        logP = log(P);
        dP = Q / P;
        y = x + 10.0;
    }
    double L = log(y);
    double t = 1.0 / y;
    double t2 = t * t;
    double sl = fma(t2, -1.0 / 156.0, 691.0 / 360360.0);
    sl = fma(t2, -sl, 1.0 / 1188.0);
    sl = fma(t2, -sl, 1.0 / 1680.0);
    sl = fma(t2, -sl, 1.0 / 1260.0);
    sl = fma(t2, -sl, 1.0 / 360.0);
    sl = fma(t2, -sl, 1.0 / 12.0);
    double sd = fma(t2, -1.0 / 12.0, 691.0 / 32760.0);
    sd = fma(t2, -sd, 1.0 / 132.0);
    sd = fma(t2, -sd, 1.0 / 240.0);
    sd = fma(t2, -sd, 1.0 / 252.0);
    sd = fma(t2, -sd, 1.0 / 120.0);
    sd = fma(t2, -sd, 1.0 / 12.0);
    lg = fma(y - 0.5, L, -y) + 0.91893853320467274178 + t * sl - logP;
    dg = L - 0.5 * t - t2 * sd - dP;
}

__device__ __forceinline__ double lgam(double x) {
    double lg, dg;
    lgam_digam(x, lg, dg);
    return lg;
}

// softplus(u) = log(1 + e^u), sigmoid(u), both from one exp + one log1p
__device__ __forceinline__ void softplus_sigmoid(double u, double& sp, double& sg) {
    double e = exp(-fabs(u));
    double l = log1p(e);
    double inv = 1.0 / (1.0 + e);
    sp = fmax(u, 0.0) + l;
    sg = (u >= 0.0) ? inv : e * inv;
}

__device__ __forceinline__ double logaddexp(double a, double b) {
    if (a == -INFINITY && b == -INFINITY) return -INFINITY;
    double m = fmax(a, b);
    return m + log(exp(a - m) + exp(b - m));
}

}  // namespace mdg
