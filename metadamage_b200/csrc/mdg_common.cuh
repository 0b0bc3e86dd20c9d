// mdg_common.cuh — device helpers shared by the sm_100a kernels: error plumbing, Philox4x32-10
// streams, warp/sub-warp reductions and the FP64 special functions (lgamma + digamma in one pass).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "mdg.h"

// Kernel tuning switches kept for A/B measurements (profiles/r01_nuts_tuning.md). Defaults = shipped.
#ifndef MDG_LOGFN
#define MDG_LOGFN 1          // 1: table-driven log_pos; 0: CUDA log()
#endif
#ifndef MDG_COLD_INLINE
#define MDG_COLD_INLINE 1    // 1: everything inlined; 0: cold helpers (Philox, Box-Muller, logaddexp, ...) out of line
#endif
#if MDG_COLD_INLINE
#define MDG_COLD __forceinline__
#else
#define MDG_COLD __noinline__
#endif

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

__device__ MDG_COLD uint4 philox4x32(uint2 key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
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



// ---------------------------------------------------------------------------------------------
// Fast path special functions of the fit kernels. Coefficients live in constant memory so that
// the FP64 instructions read them as c[bank][offset] operands (ncu on the first version showed
// 31 % of all issued instructions were moves materialising FP64 literals).
// ---------------------------------------------------------------------------------------------
__constant__ double kLogCoef[9] = {
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,
    6.93147180369123816490e-01, 1.90821492927058770002e-10};
// Stirling series of lgamma: 1/12, 1/360, 1/1260, 1/1680, 1/1188, 691/360360, 1/156 (alternating signs)
__constant__ double kStirLg[7] = {1.0 / 12.0, 1.0 / 360.0, 1.0 / 1260.0, 1.0 / 1680.0, 1.0 / 1188.0, 691.0 / 360360.0, 1.0 / 156.0};
// asymptotic series of digamma: 1/12, 1/120, 1/252, 1/240, 1/132, 691/32760, 1/12
__constant__ double kStirDg[7] = {1.0 / 12.0, 1.0 / 120.0, 1.0 / 252.0, 1.0 / 240.0, 1.0 / 132.0, 691.0 / 32760.0, 1.0 / 12.0};

// ---------------------------------------------------------------------------------------------
// Division-free natural log of a positive, normal, finite double.
//   x = 2^e * m, m in [1, 2);  i = top 7 mantissa bits;  c_i = 1 + (i + 1/2) / 128
//   r = m / c_i - 1  (|r| <= 2^-8, one fma with the tabulated 1/c_i)
//   log x = e ln2 + log c_i + (r - r^2/2 + ... - r^6/6)          (r^7/7 < 2e-18)
// The 2 KB table {1/c_i, log c_i} lives in shared memory (one copy per CTA); every kernel that
// evaluates logs calls log_table_init() once before use. Absolute error ~1e-17 (relative accuracy
// degrades only for x within 2^-8 of 1, where the result is < 4e-3 and feeds sums of O(1) terms).
// The fdlibm-style log it replaces spent 10 of its ~27 FP64 instructions in f / (2 + f).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2* log_table() {
    __shared__ double2 tab[128];
    return tab;
}

__device__ __forceinline__ double* exp_table() {
    __shared__ double tab[64];  // 2^(j/64)
    return tab;
}

__device__ __forceinline__ void log_table_init() {
    double2* tab = log_table();
    double* et = exp_table();
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        const double c = 1.0 + ((double)i + 0.5) / 128.0;
        tab[i] = make_double2(1.0 / c, log(c));
        if (i < 64) et[i] = exp2((double)i / 64.0);
    }
    __syncthreads();
}

// 1/x for positive normal x: hardware seed (rcp.approx, ~2^-23) + two Newton steps (<= 2 ulp)
__device__ __forceinline__ double rcp_pos(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// e^x for any double: x = (64 k + j) ln2/64 + r, |r| <= ln2/128; e^x = 2^k 2^(j/64) e^r with a
// degree-5 polynomial (r^6/720 < 4e-17). Results below 2^-1000 flush to 0; NaN propagates.
__device__ __forceinline__ double exp_fast(double x) {
    const double xc = fmin(fmax(x, -700.0), 709.0);
    const double n = rint(xc * 92.332482616893656877);               // 64 / ln 2
    double r = fma(n, -0.01083042469326756, xc);                     // ln2/64, high 32 bits (n * hi is exact)
    r = fma(n, -2.9815858269852933e-12, r);                          // ln2/64 - hi
    const int ni = (int)n;
    const double tj = exp_table()[ni & 63];
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(r, q, 1.0 / 6.0);
    q = fma(r, q, 0.5);
    q = fma(r, q, 1.0);
    q = fma(r, q, 1.0);
    const int k = ni >> 6;
    const double scale = __hiloint2double((k + 1023) << 20, 0);     // k in [-1010, 1023]
    const double y = (tj * q) * scale;
    return x != x ? x : (x < -700.0 ? 0.0 : (x > 709.0 ? INFINITY : y));
}

__device__ __forceinline__ double log_pos(double x) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double de = (double)((hi >> 20) - 1023);
    const double2 t = log_table()[(hi >> 13) & 127];
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double r = fma(m, t.x, -1.0);
    double q = fma(r, -1.0 / 6.0, 1.0 / 5.0);
    q = fma(r, q, -1.0 / 4.0);
    q = fma(r, q, 1.0 / 3.0);
    q = fma(r, q, -1.0 / 2.0);
    q = fma(r, q, 1.0);
    return fma(de, kLogCoef[7], t.y) + fma(de, kLogCoef[8], r * q);
}

// two standard normals (Box-Muller); sincospi needs no large-argument reduction
__device__ __forceinline__ void normal2(uint4 o, double& n0, double& n1) {
    double u0, u1;
    uniform2(o, u0, u1);
    const double rad = sqrt(-2.0 * log_pos(u0));
    double s, c;
    sincospi(2.0 * u1, &s, &c);
    n0 = rad * c;
    n1 = rad * s;
}

// Out-of-line copies of the transcendental functions for the COLD parts of the kernels (once per
// transition or rarer). The NUTS kernel was instruction-fetch bound (ncu: stall_no_instruction
// 45 % of all stalls with 12.4k SASS instructions); one shared copy of each keeps the hot loop
// (one gradient evaluation + leaf bookkeeping) inside the instruction caches.
__device__ MDG_COLD double exp_cold(double x) { return exp(x); }
__device__ MDG_COLD double log_cold(double x) { return log(x); }
__device__ MDG_COLD double sigmoid_cold(double x) { return 1.0 / (1.0 + exp(-x)); }

// st = lgamma(y) - 0.5 log(2 pi) and dg = digamma(y) for y >= 10 (7 Bernoulli terms each)
__device__ __forceinline__ double log_sel(double x) {
#if MDG_LOGFN
    return log_pos(x);
#else
    return log(x);
#endif
}

__device__ __forceinline__ void stirling(double y, double t, double& st, double& dg) {
    const double L = log_sel(y);
    const double t2 = t * t;
    double sl = fma(t2, -kStirLg[6], kStirLg[5]);
    sl = fma(t2, -sl, kStirLg[4]);
    sl = fma(t2, -sl, kStirLg[3]);
    sl = fma(t2, -sl, kStirLg[2]);
    sl = fma(t2, -sl, kStirLg[1]);
    sl = fma(t2, -sl, kStirLg[0]);
    double sd = fma(t2, -kStirDg[6], kStirDg[5]);
    sd = fma(t2, -sd, kStirDg[4]);
    sd = fma(t2, -sd, kStirDg[3]);
    sd = fma(t2, -sd, kStirDg[2]);
    sd = fma(t2, -sd, kStirDg[1]);
    sd = fma(t2, -sd, kStirDg[0]);
    st = fma(t, sl, fma(y - 0.5, L, -y));
    dg = fma(-t2, sd, fma(-0.5, t, L));
}

// lgamma(x) and digamma(x), x > 0, in one pass. x < 10 is shifted to y = x + 10 with the product
// P = x (x+1) ... (x+9) and its derivative P' (lgamma(x) = lgamma(y) - log P, digamma(x) =
// digamma(y) - P'/P: one log and one division instead of ten); y >= 10 uses the Stirling /
// asymptotic series with 7 Bernoulli terms (truncation < 4e-17 absolute at y = 10).
__device__ __forceinline__ void lgam_digam_u(double x, unsigned gmask, double& lg, double& dg) {
    double y = x, logP = 0.0, dP = 0.0;
    const bool small = x < 10.0;
    double t;
    if (small) {
        double P = x, Q = 1.0;
#pragma unroll
        for (int i = 1; i < 10; ++i) {
            const double xi = x + (double)i;
            Q = fma(Q, xi, P);
            P *= xi;
        }
        logP = log_sel(P);
        y = x + 10.0;
        const double r = rcp_pos(y * P);  // one reciprocal serves 1/y and Q/P
        t = r * P;
        dP = Q * (r * y);
    } else {
        t = rcp_pos(y);
    }
    double st, d;
    stirling(y, t, st, d);
    lg = (st + 0.91893853320467274178) - logP;
    dg = d - dP;
}

// single-thread forms (log C(N,k), the predictive's BTRS sampler, the special-function test hook)
__device__ __forceinline__ void lgam_digam(double x, double& lg, double& dg) {
    double y = x, logP = 0.0, dP = 0.0;
    if (x < 10.0) {
        double P = x, Q = 1.0;
#pragma unroll
        for (int i = 1; i < 10; ++i) {
            const double xi = x + (double)i;
            Q = fma(Q, xi, P);
            P *= xi;
        }
        logP = log_sel(P);
        dP = Q / P;
        y = x + 10.0;
    }
    double st, d;
    stirling(y, rcp_pos(y), st, d);
    lg = (st + 0.91893853320467274178) - logP;
    dg = d - dP;
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

// log(e^a + e^b) = max + log(1 + e^-|a-b|); out of line (used twice per leaf, never in the gradient)
__device__ MDG_COLD double logaddexp(double a, double b) {
    if (a == -INFINITY && b == -INFINITY) return -INFINITY;
    const double m = fmax(a, b), d = -fabs(a - b);
    return isnan(d) ? (a + b) : m + log_pos(1.0 + exp(d));
}

}  // namespace mdg
