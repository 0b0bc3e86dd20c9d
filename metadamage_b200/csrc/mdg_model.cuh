// mdg_model.cuh — K2: beta-binomial log-joint and analytic gradient, one lane per position.
//
// Follows fits.py:43-59 (model_PMD) and fits.py:62-67 (model_null) with numpyro 0.4.1's
// BetaBinomial.log_prob and the sigmoid / exp bijections of its unconstrained parameterisation
// (SURVEY.md 8c-notes). A GW-wide lane group (32 = a warp, 16 = half a warp) owns one chain;
// lane `lig` holds NPL positions; sums go through __shfl_xor_sync butterflies so that every
// lane of the group ends with bit-identical values and all control flow stays group-uniform.
#pragma once
#include "mdg_common.cuh"

namespace mdg {

struct Priors {
    double qa, qb, Aa, Ab, ca, cb;  // Beta priors (fits.py:46-48, 63)
    double rate, phi_min;           // Exponential(rate) on delta = phi - phi_min (fits.py:53-54)
    double nlb_q, nlb_A, nlb_c;     // -log B(a,b) of the three Beta priors
    double log_rate;
};

template <int MODEL>
struct ModelDim { static constexpr int value = MODEL == 0 ? 4 : 2; };

// per-lane observations; inactive slots carry k = N = 0 (their terms are masked out of the sums,
// and cost nothing: R(x, 0) = 0)
template <int NPL>
struct LaneObs {
    double k[NPL], N[NPL], x[NPL];
    bool act[NPL];
};

template <int NPL, int GW>
__device__ __forceinline__ void load_obs(LaneObs<NPL>& ob, const uint32_t* __restrict__ k,
                                         const uint32_t* __restrict__ N, int P, int mask, int lig) {
    const int n_obs = mask == 0 ? 2 * P : P;
    const int base = mask == 2 ? P : 0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        int i = s * GW + lig;
        bool a = i < n_obs;
        int dense = base + (a ? i : 0);
        ob.act[s] = a;
        ob.k[s] = a ? (double)__ldg(k + dense) : 0.0;
        ob.N[s] = a ? (double)__ldg(N + dense) : 0.0;
        ob.x[s] = a ? (double)(dense < P ? dense : dense - P) : 0.0;
    }
}

// log C(N,k): the parameter-free part of BetaBinomial.log_prob
template <int NPL>
__device__ __forceinline__ void log_binom_coeff(const LaneObs<NPL>& ob, double (&logC)[NPL]) {
#pragma unroll
    for (int s = 0; s < NPL; ++s)
        logC[s] = ob.act[s] ? lgam(ob.N[s] + 1.0) - lgam(ob.k[s] + 1.0) - lgam(ob.N[s] - ob.k[s] + 1.0) : 0.0;
}

// Per-lane coefficients of the prior (+ Jacobian) term: lane j < D owns parameter j and contributes
//   c0 + c1 u + c2 softplus(u) + c3 e^u
// to the log density, hence c1 + c2 sigmoid(u) + c3 e^u to d/du_j:
//   Beta(a, b) on sigmoid(u):  log p = (a-1+jac)(u - sp) - (b-1+jac) sp - log B(a,b)
//                              -> c0 = -log B, c1 = a-1+jac, c2 = -(a+b-2+2 jac), c3 = 0
//   Exponential(rate) on e^u:  log p = log rate - rate e^u + jac u
//                              -> c0 = log rate, c1 = jac, c2 = 0, c3 = -rate
// Lanes >= D carry zeros. The table ([2][32] double2, conflict-free 128-bit loads) lives in shared
// memory and is filled once per CTA by prior_table_init.
template <int MODEL>
__device__ __forceinline__ void prior_table_init(double2* tab, const Priors& pr, int jac) {
    constexpr int D = ModelDim<MODEL>::value;
    const double dj = (double)jac;
    for (int l = threadIdx.x; l < 32; l += blockDim.x) {
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
        if (l < D - 1) {
            double a = pr.qa, b = pr.qb, nlb = pr.nlb_q;
            if (MODEL == 0 && l == 1) { a = pr.Aa; b = pr.Ab; nlb = pr.nlb_A; }
            if (MODEL == 0 && l == 2) { a = pr.ca; b = pr.cb; nlb = pr.nlb_c; }
            c0 = nlb; c1 = a - 1.0 + dj; c2 = -((a - 1.0 + dj) + (b - 1.0 + dj));
        } else if (l == D - 1) {
            c0 = pr.log_rate; c1 = dj; c3 = -pr.rate;
        }
        tab[l] = make_double2(c0, c1);
        tab[32 + l] = make_double2(c2, c3);
    }
    __syncthreads();
}

// Evaluate log p(y | theta(u)) + log prior(theta(u)) [+ log |d theta / d u| if the prior table was
// built with jac] and its gradient w.r.t. u, for the group's chain. `ll[s]` is the per-position
// log-likelihood WITHOUT log C(N,k). `valid` is false where the reference would produce NaN
// (clip(Dz,0,1) reaching 1, fits.py:50), where |u_j| >= 700 (e^u or the sigmoid saturate; also
// catches NaN/inf states) or anything is non-finite. `has_spare`: the group's last slot is
// inactive (k = N = 0), so its evaluations are exactly the position-independent lgamma(phi) [null:
// also lgamma(alpha), lgamma(beta)] that every lane needs: they are broadcast instead of being
// computed again.
template <int MODEL, int NPL, int GW>
__device__ __forceinline__ void eval_model(const LaneObs<NPL>& ob, const double (&u)[ModelDim<MODEL>::value],
                                           const double2* __restrict__ ptab, double phi_min, bool has_spare,
                                           unsigned gmask, int lig, double& logp,
                                           double (&grad)[ModelDim<MODEL>::value], double (&ll)[NPL], bool& valid) {
    constexpr int D = ModelDim<MODEL>::value;

    // --- group-uniform transforms, one parameter per lane, then broadcast --------------------
    double myu = u[0];
#pragma unroll
    for (int j = 1; j < D; ++j) myu = (lig == j) ? u[j] : myu;
    const bool in_range = fabs(myu) < 700.0;  // false for NaN
    // ONE exponential per lane: e^-|u| for the sigmoid lanes, e^u for the phi lane
    const double E = exp_core(lig == D - 1 ? myu : -fabs(myu));
    const double l1p = log_pos(1.0 + E);
    const double inv = rcp_pos(1.0 + E);
    const double sp = fmax(myu, 0.0) + l1p;             // softplus(u)
    const double sg = (myu >= 0.0) ? inv : E * inv;     // sigmoid(u)
    // this lane's prior (+ Jacobian) term of the log density and of the gradient
    const double2 c01 = ptab[lig], c23 = ptab[32 + lig];
    const double lp_lane = fma(c23.y, E, fma(c23.x, sp, fma(c01.y, myu, c01.x)));
    const double gp_lane = fma(c23.y, E, fma(c23.x, sg, c01.y));
    const double q = __shfl_sync(gmask, sg, 0, GW);
    const double log1mq = -__shfl_sync(gmask, sp, 0, GW);
    const double delta = __shfl_sync(gmask, E, D - 1, GW);
    double A = 0.0, c = 0.0;
    if (MODEL == 0) {
        A = __shfl_sync(gmask, sg, 1, GW);
        c = __shfl_sync(gmask, sg, 2, GW);
    }
    const double phi = delta + phi_min;

    // --- per-position special functions -------------------------------------------------------
    double lg1[NPL], lg2[NPL], lg3[NPL], dg1[NPL], dg2[NPL], dg3[NPL];
    double lga[NPL], lgb[NPL], dga[NPL], dgb[NPL], Dz[NPL], w[NPL];
    bool bad = !in_range;
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        w[s] = (MODEL == 0) ? exp_nonpos(ob.x[s] * log1mq) : 1.0;
        double Dv = (MODEL == 0) ? fma(A, w[s], c) : q;
        bool ok = (Dv > 0.0) && (Dv < 1.0);
        bad |= (!ok) && ob.act[s];
        Dv = ok ? Dv : 0.5;
        Dz[s] = Dv;
        double al = Dv * phi, be = (1.0 - Dv) * phi;
        if (MODEL == 0) {
            const double xs[5] = {ob.k[s] + al, ob.N[s] - ob.k[s] + be, ob.N[s] + phi, al, be};
            double l5[5], d5[5];
            lgam_digam_batch<5>(xs, gmask, l5, d5);
            lg1[s] = l5[0]; lg2[s] = l5[1]; lg3[s] = l5[2]; lga[s] = l5[3]; lgb[s] = l5[4];
            dg1[s] = d5[0]; dg2[s] = d5[1]; dg3[s] = d5[2]; dga[s] = d5[3]; dgb[s] = d5[4];
        } else {
            const double xs[3] = {ob.k[s] + al, ob.N[s] - ob.k[s] + be, ob.N[s] + phi};
            double l3[3], d3[3];
            lgam_digam_batch<3>(xs, gmask, l3, d3);
            lg1[s] = l3[0]; lg2[s] = l3[1]; lg3[s] = l3[2];
            dg1[s] = d3[0]; dg2[s] = d3[1]; dg3[s] = d3[2];
        }
    }
    // position-independent pieces: the spare slot (k = N = 0) has just evaluated them
    double lgphi, dgphi;
    has_spare = (GW == 32) ? (__all_sync(0xffffffffu, has_spare) != 0) : has_spare;  // provably uniform: plain branches
    if (has_spare) {
        lgphi = __shfl_sync(gmask, lg3[NPL - 1], GW - 1, GW);
        dgphi = __shfl_sync(gmask, dg3[NPL - 1], GW - 1, GW);
    } else {
        lgam_digam_u(phi, gmask, lgphi, dgphi);
    }
    if (MODEL == 1) {
        double ua, da_, ub, db_;
        if (has_spare) {
            ua = __shfl_sync(gmask, lg1[NPL - 1], GW - 1, GW);
            da_ = __shfl_sync(gmask, dg1[NPL - 1], GW - 1, GW);
            ub = __shfl_sync(gmask, lg2[NPL - 1], GW - 1, GW);
            db_ = __shfl_sync(gmask, dg2[NPL - 1], GW - 1, GW);
        } else {
            lgam_digam_u(q * phi, gmask, ua, da_);
            lgam_digam_u((1.0 - q) * phi, gmask, ub, db_);
        }
#pragma unroll
        for (int s = 0; s < NPL; ++s) { lga[s] = ua; dga[s] = da_; lgb[s] = ub; dgb[s] = db_; }
    }

    // --- per-lane partial sums ------------------------------------------------------------------
    double s_ll = lp_lane, s_dD = 0.0, s_dDw = 0.0, s_dDxw = 0.0, s_dphi = 0.0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        double lls = lg1[s] + lg2[s] - lg3[s] - lga[s] - lgb[s] + lgphi;
        double dgN = dg3[s] - dgphi;
        double ga = dg1[s] - dga[s] - dgN;
        double gb = dg2[s] - dgb[s] - dgN;
        double dD = phi * (ga - gb);
        double dphi = fma(Dz[s], ga - gb, gb);  // D*ga + (1-D)*gb
        ll[s] = lls;
        if (ob.act[s]) {
            s_ll += lls;
            s_dD += dD;
            s_dphi += dphi;
            if (MODEL == 0) {
                double dw = dD * w[s];
                s_dDw += dw;
                s_dDxw = fma(dw, ob.x[s], s_dDxw);
            }
        }
    }
    s_ll = group_sum<GW>(s_ll, gmask);
    s_dD = group_sum<GW>(s_dD, gmask);
    s_dphi = group_sum<GW>(s_dphi, gmask);
    if (MODEL == 0) {
        s_dDw = group_sum<GW>(s_dDw, gmask);
        s_dDxw = group_sum<GW>(s_dDxw, gmask);
    }
    const bool any_bad = (__ballot_sync(gmask, bad) & gmask) != 0u;

    // --- chain rule to the unconstrained parameters ----------------------------------------------
    logp = s_ll;
    if (MODEL == 0) {
        grad[0] = fma(-A * q, s_dDxw, __shfl_sync(gmask, gp_lane, 0, GW));
        grad[1] = fma(A * (1.0 - A), s_dDw, __shfl_sync(gmask, gp_lane, 1, GW));
        grad[2] = fma(c * (1.0 - c), s_dD, __shfl_sync(gmask, gp_lane, 2, GW));
        grad[3] = fma(delta, s_dphi, __shfl_sync(gmask, gp_lane, 3, GW));
    } else {
        grad[0] = fma(q * (1.0 - q), s_dD, __shfl_sync(gmask, gp_lane, 0, GW));
        grad[1] = fma(delta, s_dphi, __shfl_sync(gmask, gp_lane, 1, GW));
    }
    bool fin = isfinite(logp);
#pragma unroll
    for (int j = 0; j < D; ++j) fin = fin && isfinite(grad[j]);
    valid = fin && !any_bad;
}

// constrained parameters (q, A, c, phi) of an unconstrained state; null model: A = c = NaN
template <int MODEL>
__device__ __forceinline__ void constrain(const double (&u)[ModelDim<MODEL>::value], double phi_min, double (&th)[4]) {
    th[0] = sigmoid_cold(u[0]);
    if (MODEL == 0) {
        th[1] = sigmoid_cold(u[1]);
        th[2] = sigmoid_cold(u[2]);
        th[3] = exp_cold(u[3]) + phi_min;
    } else {
        th[1] = nan("");
        th[2] = nan("");
        th[3] = exp_cold(u[1]) + phi_min;
    }
}

}  // namespace mdg
