// mdg_fit_kernels.cuh — K3 (MAP), K4 (NUTS), K5 (WAIC accumulation) kernels.
//
// One lane group (a warp, or half a warp when forward and reverse chains are packed) owns one
// (TaxID, run) work item and pulls items from an atomic counter (persistent CTAs). The NUTS
// transition restates numpyro 0.4.1 (hmc.py / hmc_util.py: build_tree, _iterative_build_subtree,
// _combine_tree, _is_turning, warmup_adapter, dual_averaging, welford_covariance,
// find_reasonable_step_size) as driven by fits.py:382-387 with the kwargs of fits.py:792-799.
// It is written as a state machine in which EVERY loop trip performs exactly one leapfrog (one
// log-density-gradient evaluation): initial-point search, step-size heuristic and tree leaves
// are just different bookkeeping after the same evaluation. That keeps a single inlined copy of
// the special-function code per model and lets two 16-lane groups with different tree shapes
// share a warp without serialising the expensive part.
#pragma once
#include "mdg_model.cuh"

namespace mdg {

#ifndef MDG_PARK
#define MDG_PARK 1
#endif
constexpr int kMaxTreeDepth = 10;
constexpr int kMaxWindows = 16;

struct RunRecord {
    double step_size, mean_accept;
    uint32_t n_leapfrog, n_divergent;
    double waic, lppd;
    double mean[5], sd[5];  // q, phi, A+c, A, c (posterior mean / std, ddof 0)
    uint32_t failed, pad;
};

struct FitLaunch {
    const int64_t* tax_id;
    const uint32_t* k;
    const uint32_t* N;
    int n_tax, P;
    mdg_fit_config cfg;
    Priors pr;
    int n_windows;
    int win_end[kMaxWindows];
    unsigned int* work_counter;
    int n_items;
    int n_masks;        // GW=32 launches: masks handled per TaxID (1: all; 2: fwd,rev; 3: all,fwd,rev)
    int mask0;          // first mask of this launch
    RunRecord* rec;     // [n_tax][6]
    double* waic;       // [n_tax][6][2][2P]: lppd_i, pWAIC_i
    double* samples;    // [n_tax][sample_runs][S][4] constrained draws, or NULL
    int sample_slot[MDG_NUM_RUNS];  // run kind -> slot in `samples`, -1 = not stored
    int sample_runs;
    double* trace;      // [n_tax][6][W+S][4] or NULL
    int n_slots;        // group kernel: rounds of the position loop, ceil((n_obs + 1) / GW)
    double* waic_acc;   // group kernel: WAIC accumulators, [grid * warps][4][n_slots][32]
};

template <int D>
struct GroupShared {
    double zl[D], rl[D], gl[D], zr[D], rr[D], gr[D];  // main tree edges
    double zp[D], gp[D];                              // main proposal == chain state between transitions
    double szp[D], sgp[D];                            // subtree proposal
    double rck[kMaxTreeDepth][D], rsck[kMaxTreeDepth][D];
    double wf_mean[D], wf_m2[D];
    double acc_mean[5], acc_m2[5];
    // cold group-uniform chain state (touched once per transition / doubling, never inside the
    // gradient evaluation). Kept here, 8 bytes per value per chain, instead of in registers: the
    // compiler's spills would replicate each of them 32x in local memory (ncu: 866 B of spills per
    // thread at 128 registers = 443 KB per SM, more than the L1 holds).
    double da_x, da_xavg, da_gavg, da_prox, mean_accept, h_step, h_Er;
    double m_weight, m_sum_acc, m_pe_p, u_main, pe_cur, eps, s_pe_p;
    double m_rsum[D];
#if MDG_PARK
    // group-uniform values that are live ACROSS the gradient evaluation but not used inside it:
    // parked here so that the evaluation's five interleaved special-function chains get the registers
    double imm[D], s_rsum[D], zn_park[D], rh_park[D], E0, s_weight, s_sum_acc;
#endif
    double isd[D];  // group kernel: 1 / sqrt(imm) = sqrt of the mass matrix diagonal (momentum r ~ N(0, M) is n * isd)
    int da_t, wf_n, window_idx, h_last, h_dir, m_nprop;
    uint32_t h_att, h_call, init_attempt, n_div;
};

// cold per-lane state of one warp: element (slot s, lane l) at [s][l] (bank-conflict free)
template <int NPL>
struct WarpLanes {
    double ll_cur[NPL][32], ll_main[NPL][32], w_max[NPL][32], w_sum[NPL][32], w_mean[NPL][32], w_m2[NPL][32], logC[NPL][32];
};

struct LaneRef {  // view of one lane's column of a [NPL][32] array, indexed by slot
    double* p;
    __device__ __forceinline__ double& operator[](int s) const { return p[s * 32]; }
};

enum Phase : int { PH_INIT = 0, PH_HEUR = 1, PH_LEAF = 2 };

template <int D>
__device__ __forceinline__ double kinetic(const double (&imm)[D], const double (&r)[D]) {
    double e = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) e = fma(imm[j] * r[j], r[j], e);
    return 0.5 * e;
}

// generalised U-turn test (hmc_util._is_turning, diagonal mass matrix)
template <int D>
__device__ __forceinline__ bool is_turning(const double (&imm)[D], const double* r_left, const double* r_right,
                                           const double* r_sum) {
    double dl = 0.0, dr = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double rho = r_sum[j] - (r_left[j] + r_right[j]) / 2;
        dl += imm[j] * r_left[j] * rho;
        dr += imm[j] * r_right[j] * rho;
    }
    return (dl <= 0.0) || (dr <= 0.0);
}

struct Vec4 { double v[4]; };

// momentum r ~ N(0, M) with M^-1 = diag(imm): out of line (one copy; Box-Muller + Philox are cold)
template <int D>
__device__ MDG_COLD Vec4 draw_momentum_cold(uint2 key, uint32_t c1, uint32_t c2, uint32_t c3, Vec4 imm) {
    Vec4 r;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        double n0 = 0.0, n1 = 0.0;
        if (2 * b < D) normal2(philox4x32(key, (uint32_t)b, c1, c2, c3), n0, n1);
        r.v[2 * b] = 2 * b < D ? n0 / sqrt(imm.v[2 * b]) : 0.0;
        r.v[2 * b + 1] = 2 * b + 1 < D ? n1 / sqrt(imm.v[2 * b + 1]) : 0.0;
    }
    return r;
}

template <int D>
__device__ __forceinline__ void draw_momentum(uint2 key, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const double (&imm)[D], double (&r)[D]) {
    Vec4 m;
#pragma unroll
    for (int j = 0; j < 4; ++j) m.v[j] = j < D ? imm[j] : 1.0;
    const Vec4 out = draw_momentum_cold<D>(key, c1, c2, c3, m);
#pragma unroll
    for (int j = 0; j < D; ++j) r[j] = out.v[j];
}

// ---------------------------------------------------------------------------------------------
// K4: NUTS. MODEL 0 = PMD, 1 = null. NPL positions per lane. GW lanes per chain.
// ---------------------------------------------------------------------------------------------
#ifndef MDG_NUTS_MINBLOCKS
#define MDG_NUTS_MINBLOCKS 4
#endif
// SPARE: every run of the launch leaves the group's last slot without a position (n_obs < NPL * GW), so the
// position-independent lgamma/digamma values come from that slot (eval_model). A template parameter:
// the fallback (three more inlined evaluations) then does not exist in the hot loop's code at all.
template <int MODEL, int NPL, int GW, int WARPS, bool SPARE>
__global__ void __launch_bounds__(WARPS * 32, MDG_NUTS_MINBLOCKS) nuts_kernel(const FitLaunch p) {
    constexpr int D = ModelDim<MODEL>::value;
    constexpr int GROUPS = 32 / GW;
    __shared__ GroupShared<D> sh_all[WARPS * GROUPS];
    __shared__ WarpLanes<NPL> sh_lanes[WARPS];
    __shared__ unsigned int sh_item[WARPS];
    __shared__ double2 sh_prior[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / GW, lig = lane % GW;
    const unsigned gmask = group_mask<GW>();
    GroupShared<D>& sh = sh_all[warp * GROUPS + grp];
    const int W = p.cfg.num_warmup, S = p.cfg.num_samples, P = p.P;
    const int max_depth = p.cfg.max_tree_depth < kMaxTreeDepth ? p.cfg.max_tree_depth : kMaxTreeDepth;
    const double log_target_heur = -0.22314355131420976;  // log(0.8)
    log_table_init();
    prior_table_init<MODEL>(sh_prior, p.pr, 1);
    const double phi_min = p.pr.phi_min;
    const uint32_t budget = p.cfg.max_leapfrogs_per_run > 0 ? (uint32_t)p.cfg.max_leapfrogs_per_run : 0u;

    for (;;) {
        if (lane == 0) sh_item[warp] = atomicAdd(p.work_counter, 1u);
        __syncwarp();
        const unsigned item = sh_item[warp];
        __syncwarp();
        if (item >= (unsigned)p.n_items) break;
        int tax, mask;
        if (GW == 32) { tax = (int)(item / (unsigned)p.n_masks); mask = p.mask0 + (int)(item % (unsigned)p.n_masks); }
        else { tax = (int)item; mask = 1 + grp; }
        const int run_kind = mask * 2 + MODEL;
        const bool has_spare = SPARE;  // the host checked (mask == 0 ? 2 * P : P) < NPL * GW for every mask of the launch
        const uint2 key = make_key(p.cfg.seed, p.tax_id[tax]);

        LaneObs<NPL> ob;
        load_obs<NPL, GW>(ob, p.k + (size_t)tax * 2 * P, p.N + (size_t)tax * 2 * P, P, mask, lig);
        WarpLanes<NPL>& wl = sh_lanes[warp];
        const LaneRef logC{&wl.logC[0][lane]}, ll_cur{&wl.ll_cur[0][lane]}, ll_main{&wl.ll_main[0][lane]};
        const LaneRef w_max{&wl.w_max[0][lane]}, w_sum{&wl.w_sum[0][lane]}, w_mean{&wl.w_mean[0][lane]}, w_m2{&wl.w_m2[0][lane]};
        {
            double lc[NPL];
            log_binom_coeff<NPL>(ob, lc);
#pragma unroll
            for (int s = 0; s < NPL; ++s) {
                logC[s] = lc[s];
                w_max[s] = -INFINITY; w_sum[s] = 0.0; w_mean[s] = 0.0; w_m2[s] = 0.0; ll_cur[s] = 0.0; ll_main[s] = 0.0;
            }
        }

        // ---- hot chain state in registers (group-uniform unless noted) ----
#if MDG_PARK
        double (&imm)[D] = sh.imm;
#else
        double imm[D];
#endif
#pragma unroll
        for (int j = 0; j < D; ++j) imm[j] = 1.0;
        double ll_sub[NPL];  // per lane
#pragma unroll
        for (int s = 0; s < NPL; ++s) ll_sub[s] = 0.0;
        uint32_t n_grad = 0;
        int failed = 0;
#if MDG_PARK
        double& E0 = sh.E0; double (&s_rsum)[D] = sh.s_rsum; double& s_weight = sh.s_weight; double& s_sum_acc = sh.s_sum_acc;
        E0 = 0.0; s_weight = 0.0; s_sum_acc = 0.0;
#else
        double E0 = 0.0;
        double s_rsum[D];
        double s_weight = 0.0, s_sum_acc = 0.0;
#endif
        int m_depth = 0;
        bool m_turning = false, m_div = false, going_right = true;
        int s_nprop = 0;
        bool s_div = false;
        uint32_t leaf_counter = 0;
        // leapfrog source
        double zf[D], rf[D], gf[D], e = 0.0;
        int phase = PH_INIT;
        int t = 0;
        // ---- cold chain state: references into shared memory (every lane of the group writes
        // the same value, reads are broadcasts) ----
        double& eps = sh.eps; double& pe_cur = sh.pe_cur;
        double& da_x = sh.da_x; double& da_xavg = sh.da_xavg; double& da_gavg = sh.da_gavg; double& da_prox = sh.da_prox;
        int& da_t = sh.da_t; int& wf_n = sh.wf_n; int& window_idx = sh.window_idx;
        double& mean_accept = sh.mean_accept; uint32_t& n_div = sh.n_div;
        double& m_weight = sh.m_weight; double& m_sum_acc = sh.m_sum_acc; double& m_pe_p = sh.m_pe_p; double& u_main = sh.u_main;
        double (&m_rsum)[D] = sh.m_rsum; int& m_nprop = sh.m_nprop; double& s_pe_p = sh.s_pe_p;
        double& h_step = sh.h_step; double& h_Er = sh.h_Er; int& h_last = sh.h_last; int& h_dir = sh.h_dir;
        uint32_t& h_att = sh.h_att; uint32_t& h_call = sh.h_call; uint32_t& init_attempt = sh.init_attempt;
        __syncwarp(gmask);
        eps = p.cfg.init_step_size; pe_cur = 0.0;
        da_x = 0.0; da_xavg = 0.0; da_gavg = 0.0; da_prox = 0.0; da_t = 0; wf_n = 0; window_idx = 0;
        mean_accept = 0.0; n_div = 0u;
        m_weight = 0.0; m_sum_acc = 0.0; m_pe_p = 0.0; u_main = 0.0; m_nprop = 0; s_pe_p = 0.0;
        h_step = 0.0; h_Er = 0.0; h_last = 0; h_dir = 0; h_att = 0u; h_call = 0u; init_attempt = 0u;
        if (lig == 0) {
#pragma unroll
            for (int j = 0; j < D; ++j) { sh.wf_mean[j] = 0.0; sh.wf_m2[j] = 0.0; }
#pragma unroll
            for (int j = 0; j < 5; ++j) { sh.acc_mean[j] = 0.0; sh.acc_m2[j] = 0.0; }
        }
        __syncwarp(gmask);

        auto init_candidate = [&]() {
#pragma unroll
            for (int b = 0; b < (D + 1) / 2; ++b) {
                double u0, u1;
                uniform2(philox4x32(key, (uint32_t)b, init_attempt, c2word(run_kind, P_INIT), 0u), u0, u1);
                zf[2 * b] = p.cfg.init_radius * (2.0 * u0 - 1.0);
                if (2 * b + 1 < D) zf[2 * b + 1] = p.cfg.init_radius * (2.0 * u1 - 1.0);
            }
#pragma unroll
            for (int j = 0; j < D; ++j) { rf[j] = 0.0; gf[j] = 0.0; }
            e = 0.0;
            phase = PH_INIT;
        };

        // start one tree doubling: pick direction, aim the next leapfrog at the chosen edge
        auto start_doubling = [&]() {
            double u_dir;
            uniform2(philox4x32(key, (uint32_t)m_depth, (uint32_t)t, c2word(run_kind, P_DIR), 0u), u_dir, u_main);
            going_right = u_dir < 0.5;
            s_nprop = 0;
            __syncwarp(gmask);
#pragma unroll
            for (int j = 0; j < D; ++j) {
                zf[j] = going_right ? sh.zr[j] : sh.zl[j];
                rf[j] = going_right ? sh.rr[j] : sh.rl[j];
                gf[j] = going_right ? sh.gr[j] : sh.gl[j];
            }
            e = going_right ? eps : -eps;
            phase = PH_LEAF;
        };

        // begin a transition from the chain state held in sh.zp / sh.gp / pe_cur
        auto start_transition = [&]() {
            double r0[D];
            draw_momentum<D>(key, (uint32_t)t, c2word(run_kind, P_MOM), 0u, imm, r0);
            E0 = pe_cur + kinetic<D>(imm, r0);
            __syncwarp(gmask);
            if (lig == 0) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    double zj = sh.zp[j], gj = sh.gp[j];
                    sh.zl[j] = zj; sh.zr[j] = zj; sh.gl[j] = gj; sh.gr[j] = gj;
                    sh.rl[j] = r0[j]; sh.rr[j] = r0[j];
                }
            }
#pragma unroll
            for (int j = 0; j < D; ++j) m_rsum[j] = r0[j];
            m_weight = 0.0; m_sum_acc = 0.0; m_nprop = 0; m_depth = 0; m_turning = false; m_div = false;
            m_pe_p = pe_cur;
#pragma unroll
            for (int s = 0; s < NPL; ++s) ll_main[s] = ll_cur[s];
            leaf_counter = 0;
            start_doubling();
        };

        // heuristic step-size search (hmc_util.find_reasonable_step_size): next trial or finish
        auto heur_try = [&]() -> bool {
            bool small_ok = (h_step > 2.2250738585072014e-308) || (h_dir >= 0);
            bool large_ok = (h_step < 1.7976931348623157e308) || (h_dir <= 0);
            if (!(small_ok && large_ok && (h_last == 0 || h_dir == h_last))) return false;
            h_step *= (h_dir > 0 ? 2.0 : (h_dir < 0 ? 0.5 : 1.0));
            draw_momentum<D>(key, h_call, c2word(run_kind, P_HEUR), h_att, imm, rf);
            ++h_att;
            h_Er = kinetic<D>(imm, rf) + pe_cur;
            __syncwarp(gmask);
#pragma unroll
            for (int j = 0; j < D; ++j) { zf[j] = sh.zp[j]; gf[j] = sh.gp[j]; }
            e = h_step;
            phase = PH_HEUR;
            return true;
        };
        auto begin_heuristic = [&]() -> bool {
            h_step = eps; h_last = 0; h_dir = 0; h_att = 0;
            return heur_try();
        };
        auto reset_dual_averaging = [&]() {
            da_x = 0.0; da_xavg = 0.0; da_gavg = 0.0; da_t = 0; da_prox = log_cold(10.0 * eps);
        };

        init_candidate();

        // =========================== one leapfrog per trip ===========================
        for (;;) {
            double zn[D], rn[D], gn[D], pen;
            {
                double rh[D];
#pragma unroll
                for (int j = 0; j < D; ++j) { rh[j] = fma(-0.5 * e, gf[j], rf[j]); zn[j] = fma(e * imm[j], rh[j], zf[j]); }
                double logp, grad[D], ll_leaf[NPL];
                bool valid;
#if MDG_PARK
#pragma unroll
                for (int j = 0; j < D; ++j) { sh.zn_park[j] = zn[j]; sh.rh_park[j] = rh[j]; }
#endif
                eval_model<MODEL, NPL, GW>(ob, zn, sh_prior, phi_min, has_spare, gmask, lig, logp, grad, ll_leaf, valid);
#if MDG_PARK
                __syncwarp(gmask);  // (compiler barrier: the parked values are re-read, not kept in registers)
#pragma unroll
                for (int j = 0; j < D; ++j) { zn[j] = sh.zn_park[j]; rh[j] = sh.rh_park[j]; }
#endif
                ++n_grad;
                // bounded work per run: the analogue of the reference's per-fit timeout (fits.py:472-474)
                if (budget != 0u && n_grad > budget) { failed = 2; break; }
                pen = valid ? -logp : nan("");
#pragma unroll
                for (int j = 0; j < D; ++j) { gn[j] = valid ? -grad[j] : nan(""); rn[j] = fma(-0.5 * e, gn[j], rh[j]); }

                if (phase == PH_LEAF) {
                    // ---- hmc_util._build_basetree ----
                    double dE = pen + kinetic<D>(imm, rn) - E0;
                    if (isnan(dE)) dE = INFINITY;
                    const double leaf_w = -dE;
                    const bool leaf_div = dE > p.cfg.max_delta_energy;
                    const double leaf_acc = dE <= 0.0 ? 1.0 : exp_nonpos(-dE);  // min(1, e^-dE); dE is never NaN here
                    const int leaf_idx = s_nprop;
                    bool take;
                    if (leaf_idx == 0) {
                        take = true;
                        s_weight = leaf_w;
                        s_sum_acc = leaf_acc;
#pragma unroll
                        for (int j = 0; j < D; ++j) s_rsum[j] = rn[j];
                    } else {
                        // ---- _combine_tree(..., biased_transition=False) ----
                        double us, unused;
                        uniform2(philox4x32(key, leaf_counter, (uint32_t)t, c2word(run_kind, P_SUB), 0u), us, unused);
                        // expit(d) and logaddexp share one exponential: e = exp(-|d|)
                        const double dlt = leaf_w - s_weight;
                        const double ed = exp_nonpos(-fabs(dlt));  // NaN -> ~0 -> take = false
                        const double inv = rcp_pos(1.0 + ed);
                        const double prob = dlt >= 0.0 ? inv : ed * inv;
                        take = us < prob;  // NaN -> false
                        s_weight = isnan(dlt) ? -INFINITY : fmax(s_weight, leaf_w) + log_pos(1.0 + ed);
                        s_sum_acc += leaf_acc;
#pragma unroll
                        for (int j = 0; j < D; ++j) s_rsum[j] += rn[j];
                    }
                    s_div = leaf_div;
                    s_nprop = leaf_idx + 1;
                    ++leaf_counter;
                    if (take) {
                        s_pe_p = pen;
#pragma unroll
                        for (int s = 0; s < NPL; ++s) ll_sub[s] = ll_leaf[s];
                    }
                    // checkpoint indices (_leaf_idx_to_ckpt_idxs)
                    const int idx_max = __popc((unsigned)leaf_idx >> 1);
                    const int n_trail = __ffs(~(unsigned)leaf_idx) - 1;
                    const int idx_min = idx_max - n_trail + 1;
                    __syncwarp(gmask);
                    if (lig == 0) {
                        if (take) {
#pragma unroll
                            for (int j = 0; j < D; ++j) { sh.szp[j] = zn[j]; sh.sgp[j] = gn[j]; }
                        }
                        if ((leaf_idx & 1) == 0) {
#pragma unroll
                            for (int j = 0; j < D; ++j) { sh.rck[idx_max][j] = rn[j]; sh.rsck[idx_max][j] = s_rsum[j]; }
                        }
                    }
                    __syncwarp(gmask);
                    bool turning = false;
                    if (leaf_idx & 1) {
                        // ---- _is_iterative_turning ----
                        for (int i = idx_max; i >= idx_min && !turning; --i) {
                            double sub[D], rc[D];
#pragma unroll
                            for (int j = 0; j < D; ++j) { rc[j] = sh.rck[i][j]; sub[j] = s_rsum[j] - sh.rsck[i][j] + rc[j]; }
                            turning = is_turning<D>(imm, rc, rn, sub);
                        }
                    }
                    if (s_nprop < (1 << m_depth) && !turning && !s_div) {
                        // keep extending the subtree from the leaf just built
#pragma unroll
                        for (int j = 0; j < D; ++j) { zf[j] = zn[j]; rf[j] = rn[j]; gf[j] = gn[j]; }
                    } else {
                        // ---- subtree finished: _combine_tree(..., biased_transition=True) ----
                        const double dlt_m = s_weight - m_weight;
                        // min(1, e^d) and logaddexp share one exponential e^-|d| (u_main < 1, so a
                        // probability above 1 acts as 1; NaN -> ~0 -> false)
                        const double em = exp_nonpos(-fabs(dlt_m));
                        const double prob = (turning || s_div) ? 0.0 : (dlt_m >= 0.0 ? 1.0 : em);
                        const bool take_main = u_main < prob;
                        __syncwarp(gmask);
                        if (lig == 0) {
#pragma unroll
                            for (int j = 0; j < D; ++j) {
                                if (going_right) { sh.zr[j] = zn[j]; sh.rr[j] = rn[j]; sh.gr[j] = gn[j]; }
                                else { sh.zl[j] = zn[j]; sh.rl[j] = rn[j]; sh.gl[j] = gn[j]; }
                                if (take_main) { sh.zp[j] = sh.szp[j]; sh.gp[j] = sh.sgp[j]; }
                            }
                        }
                        __syncwarp(gmask);  // lane 0 is back before anybody touches group-uniform state again
#pragma unroll
                        for (int j = 0; j < D; ++j) m_rsum[j] += s_rsum[j];
                        __syncwarp(gmask);
                        {
                            double rl[D], rr[D];
#pragma unroll
                            for (int j = 0; j < D; ++j) { rl[j] = sh.rl[j]; rr[j] = sh.rr[j]; }
                            m_turning = turning || is_turning<D>(imm, rl, rr, m_rsum);
                        }
                        if (take_main) {
                            m_pe_p = s_pe_p;
#pragma unroll
                            for (int s = 0; s < NPL; ++s) ll_main[s] = ll_sub[s];
                        }
                        m_depth += 1;
                        // logaddexp(m_weight, s_weight) = max + log(1 + e^-|d|)
                        m_weight = isnan(dlt_m) ? -INFINITY : fmax(m_weight, s_weight) + log_pos(1.0 + em);
                        m_div = s_div;
                        m_sum_acc += s_sum_acc;
                        m_nprop += s_nprop;
                        if (m_depth < max_depth && !m_turning && !m_div) {
                            start_doubling();
                        } else {
                            // ================= transition finished (hmc.py sample_kernel) =================
                            const double accept_prob = m_sum_acc / (double)m_nprop;
                            pe_cur = m_pe_p;
#pragma unroll
                            for (int s = 0; s < NPL; ++s) ll_cur[s] = ll_main[s];
                            __syncwarp(gmask);
                            double zc[D];
#pragma unroll
                            for (int j = 0; j < D; ++j) zc[j] = sh.zp[j];
                            bool want_heur = false;
                            if (t < W) {
                                // ---- warmup_adapter.update_fn ----
                                da_t += 1;
                                da_gavg = (1.0 - 1.0 / (da_t + 10)) * da_gavg + (p.cfg.target_accept - accept_prob) / (da_t + 10);
                                da_x = da_prox - sqrt((double)da_t) / 0.05 * da_gavg;
                                const double wt = exp_cold(-0.75 * log_cold((double)da_t));
                                da_xavg = (1.0 - wt) * da_xavg + wt * da_x;
                                eps = exp_cold((t == W - 1) ? da_xavg : da_x);
                                eps = fmax(eps, 2.2250738585072014e-308);
                                const bool is_middle = (0 < window_idx) && (window_idx < p.n_windows - 1);
                                if (is_middle) {
                                    wf_n += 1;
                                    if (lig == 0) {
#pragma unroll
                                        for (int j = 0; j < D; ++j) {
                                            double dpre = zc[j] - sh.wf_mean[j];
                                            double mn = sh.wf_mean[j] + dpre / wf_n;
                                            sh.wf_mean[j] = mn;
                                            sh.wf_m2[j] += dpre * (zc[j] - mn);
                                        }
                                    }
                                    __syncwarp(gmask);
                                }
                                const bool at_end = (t == p.win_end[window_idx]);
                                __syncwarp(gmask);  // every lane has read window_idx before it moves
                                if (at_end) window_idx += 1;
                                if (at_end && is_middle) {
                                    __syncwarp(gmask);
#pragma unroll
                                    for (int j = 0; j < D; ++j) {
                                        double cov = sh.wf_m2[j] / (wf_n - 1);
                                        imm[j] = ((double)wf_n / (wf_n + 5.0)) * cov + 1e-3 * (5.0 / (wf_n + 5.0));
                                    }
                                    __syncwarp(gmask);
                                    if (lig == 0) {
#pragma unroll
                                        for (int j = 0; j < D; ++j) { sh.wf_mean[j] = 0.0; sh.wf_m2[j] = 0.0; }
                                    }
                                    __syncwarp(gmask);
                                    wf_n = 0;
                                    want_heur = p.cfg.find_heuristic_step_size != 0;
                                    if (!want_heur) reset_dual_averaging();
                                }
                            } else {
                                // ---- after warm-up: keep the draw ----
                                const int si = t - W;
                                mean_accept += (accept_prob - mean_accept) / (double)(si + 1);
                                if (m_div) ++n_div;
                                double th[4];
                                constrain<MODEL>(zc, p.pr.phi_min, th);
                                if (lig == 0) {
                                    const double v[5] = {th[0], th[3], MODEL == 0 ? th[1] + th[2] : th[0], th[1], th[2]};
#pragma unroll
                                    for (int j = 0; j < 5; ++j) {
                                        double dpre = v[j] - sh.acc_mean[j];
                                        double mn = sh.acc_mean[j] + dpre / (double)(si + 1);
                                        sh.acc_mean[j] = mn;
                                        sh.acc_m2[j] += dpre * (v[j] - mn);
                                    }
                                    const int slot = p.sample_slot[run_kind];
                                    if (p.samples != nullptr && slot >= 0) {
                                        double* dst = p.samples + (((size_t)tax * p.sample_runs + slot) * S + si) * 4;
                                        dst[0] = th[0]; dst[1] = th[1]; dst[2] = th[2]; dst[3] = th[3];
                                    }
                                }
                                __syncwarp(gmask);
                                // WAIC: streaming logsumexp + Welford of this lane's log-likelihood (fits.py:147-165)
#pragma unroll
                                for (int s = 0; s < NPL; ++s) {
                                    const double v = ll_cur[s];
                                    const double ed = exp_cold(-fabs(v - w_max[s]));  // exp(-inf) = 0 on the first draw
                                    if (v > w_max[s]) { w_sum[s] = fma(w_sum[s], ed, 1.0); w_max[s] = v; }
                                    else w_sum[s] += ed;
                                    const double dpre = v - w_mean[s];
                                    w_mean[s] += dpre / (double)(si + 1);
                                    w_m2[s] += dpre * (v - w_mean[s]);
                                }
                            }
                            if (p.trace != nullptr && lig == 0) {
                                double* dst = p.trace + (((size_t)tax * MDG_NUM_RUNS + run_kind) * (W + S) + t) * 4;
#pragma unroll
                                for (int j = 0; j < 4; ++j) dst[j] = j < D ? zc[j] : nan("");
                            }
                            __syncwarp(gmask);
                            t += 1;
                            if (t >= W + S) break;
                            bool heur_running = false;
                            if (want_heur) {
                                ++h_call;
                                heur_running = begin_heuristic();
                                if (!heur_running) { eps = h_step; reset_dual_averaging(); }
                            }
                            if (!heur_running) start_transition();
                        }
                    }
                } else if (phase == PH_HEUR) {
                    const double delta = (kinetic<D>(imm, rn) + pen) - h_Er;
                    const int dir_new = (log_target_heur < -delta) ? 1 : -1;  // NaN -> -1
                    h_last = h_dir;
                    h_dir = dir_new;
                    if (!heur_try()) {
                        eps = h_step;
                        reset_dual_averaging();
                        start_transition();
                    }
                } else {
                    // ---- PH_INIT: init_to_uniform(radius), retried until finite ----
                    if (valid) {
                        pe_cur = pen;
#pragma unroll
                        for (int s = 0; s < NPL; ++s) ll_cur[s] = ll_leaf[s];
                        __syncwarp(gmask);
                        if (lig == 0) {
#pragma unroll
                            for (int j = 0; j < D; ++j) { sh.zp[j] = zn[j]; sh.gp[j] = gn[j]; }
                        }
                        __syncwarp(gmask);
                        if (W + S == 0) break;
                        bool heur_running = false;
                        if (p.cfg.find_heuristic_step_size && W > 0) {
                            h_call = 0;
                            heur_running = begin_heuristic();
                            if (!heur_running) eps = h_step;
                        }
                        if (!heur_running) { reset_dual_averaging(); start_transition(); }
                    } else {
                        ++init_attempt;
                        if (init_attempt >= 100u) { failed = 1; break; }
                        init_candidate();
                    }
                }
            }
        }

        // ---- per-run outputs ----
        double waic_sum = 0.0, lppd_sum = 0.0;
        const size_t R = 2 * (size_t)P;
        double* wout = p.waic + ((size_t)tax * MDG_NUM_RUNS + run_kind) * 2 * R;
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            if (ob.act[s] && !failed && S > 0) {
                const int dense = (mask == 2 ? P : 0) + s * GW + lig;
                const double lppd_i = logC[s] + w_max[s] + log_cold(w_sum[s]) - log_cold((double)S);
                const double pw_i = w_m2[s] / (double)S;
                wout[dense] = lppd_i;
                wout[R + dense] = pw_i;
                lppd_sum += lppd_i;
                waic_sum += -2.0 * (lppd_i - pw_i);
            }
        }
        waic_sum = group_sum<GW>(waic_sum, gmask);
        lppd_sum = group_sum<GW>(lppd_sum, gmask);
        __syncwarp(gmask);
        if (lig == 0) {
            RunRecord& r = p.rec[(size_t)tax * MDG_NUM_RUNS + run_kind];
            r.step_size = eps;
            r.mean_accept = mean_accept;
            r.n_leapfrog = n_grad;
            r.n_divergent = n_div;
            r.waic = waic_sum;
            r.lppd = lppd_sum;
#pragma unroll
            for (int j = 0; j < 5; ++j) { r.mean[j] = sh.acc_mean[j]; r.sd[j] = S > 0 ? sqrt(sh.acc_m2[j] / (double)S) : 0.0; }
            r.failed = (uint32_t)failed;
            r.pad = 0;
        }
        __syncwarp(gmask);
        if (GW != 32) __syncwarp();  // both halves are done before the warp pulls the next item
    }
}

// ---------------------------------------------------------------------------------------------
// K3: MAP — mode of the constrained-space posterior density (no Jacobian term), LM-damped Newton
// in unconstrained coordinates with a central-difference Hessian of the analytic gradient.
// One warp per (TaxID, model).
// ---------------------------------------------------------------------------------------------
struct MapRecord {
    double theta[4];  // q, A, c, phi
    double logp;
    uint32_t iters, converged;
};

struct MapLaunch {
    const int64_t* tax_id;
    const uint32_t* k;
    const uint32_t* N;
    int n_tax, P;
    Priors pr;
    unsigned int* work_counter;
    MapRecord* rec;  // [n_tax][2]: PMD, null
};

template <int D>
__device__ __forceinline__ bool chol_solve(const double (&H)[D][D], const double (&rhs)[D], double (&x)[D]) {
    double L[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) L[i][j] = 0.0;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double s = H[i][j];
#pragma unroll
            for (int m = 0; m < j; ++m) s -= L[i][m] * L[j][m];
            if (i == j) { ok = ok && (s > 0.0); L[i][i] = sqrt(s); }
            else L[i][j] = s / L[j][j];
        }
    }
    if (!ok) return false;
    double y[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = rhs[i];
#pragma unroll
        for (int m = 0; m < i; ++m) s -= L[i][m] * y[m];
        y[i] = s / L[i][i];
    }
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
        double s = y[i];
#pragma unroll
        for (int m = i + 1; m < D; ++m) s -= L[m][i] * x[m];
        x[i] = s / L[i][i];
    }
    return true;
}

template <int MODEL, int NPL>
__device__ void map_fit_group(const LaneObs<NPL>& ob, const double (&logC)[NPL], int P, const Priors& pr,
                              const double2* ptab, bool has_spare, int lig, MapRecord& out) {
    constexpr int D = ModelDim<MODEL>::value;
    constexpr unsigned gmask = 0xffffffffu;
    constexpr double UMAX = 40.0;
    // ---- data-driven starting point ----
    double u[D];
    {
        double ktail = 0, Ntail = 0, k1 = 0, N1 = 0, kall = 0, Nall = 0;
        const double xmax = (double)(P - 1);
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            if (ob.act[s]) {
                kall += ob.k[s]; Nall += ob.N[s];
                if (ob.x[s] == 0.0) { k1 += ob.k[s]; N1 += ob.N[s]; }
                if (ob.x[s] >= xmax - 2.0) { ktail += ob.k[s]; Ntail += ob.N[s]; }
            }
        }
        ktail = group_sum<32>(ktail, gmask); Ntail = group_sum<32>(Ntail, gmask);
        k1 = group_sum<32>(k1, gmask); N1 = group_sum<32>(N1, gmask);
        kall = group_sum<32>(kall, gmask); Nall = group_sum<32>(Nall, gmask);
        double f_all = (kall + 0.5) / (Nall + 1.0);
        double c0 = (ktail + 0.5) / (Ntail + 1.0);
        double A0 = (k1 + 0.5) / (N1 + 1.0) - c0;
        c0 = fmin(fmax(c0, 1e-6), 0.5);
        A0 = fmin(fmax(A0, 1e-3), 0.45);
        f_all = fmin(fmax(f_all, 1e-6), 0.9);
        if (MODEL == 0) {
            u[0] = log(0.3 / 0.7);
            u[1] = log(A0 / (1.0 - A0));
            u[2] = log(c0 / (1.0 - c0));
            u[3] = log(100.0);
        } else {
            u[0] = log(f_all / (1.0 - f_all));
            u[1] = log(100.0);
        }
    }
    double f, g[D], ll[NPL], lp;
    bool ok;
    eval_model<MODEL, NPL, 32>(ob, u, ptab, pr.phi_min, has_spare, gmask, lig, lp, g, ll, ok);
    // f carries log C(N,k) so that its magnitude (and the step-acceptance slack) matches the
    // constrained-space log posterior; `slack` covers the cancellation noise of the lgamma sums
    double sumC = 0.0, scale = 0.0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) { sumC += logC[s]; scale += ob.act[s] ? lgam(ob.N[s] + 1.0) : 0.0; }
    sumC = group_sum<32>(sumC, gmask);
    scale = group_sum<32>(scale, gmask);
    const double slack = 1e-14 * scale + 1e-10;
    out.converged = 0;
    out.iters = 0;
    if (!ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) out.theta[j] = nan("");
        out.logp = nan("");
        return;
    }
    f = -(lp + sumC);
#pragma unroll
    for (int j = 0; j < D; ++j) g[j] = -g[j];
    double lambda = 1e-3;
    int it = 0;
    bool converged = false;
    for (it = 0; it < 200; ++it) {
        double H[D][D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double h = 1e-4 * (1.0 + fabs(u[j]));
            double up[D], um[D], gp[D], gm[D], t0, t1;
            bool okp, okm;
#pragma unroll
            for (int i = 0; i < D; ++i) { up[i] = u[i] + (i == j ? h : 0.0); um[i] = u[i] - (i == j ? h : 0.0); }
            eval_model<MODEL, NPL, 32>(ob, up, ptab, pr.phi_min, has_spare, gmask, lig, t0, gp, ll, okp);
            eval_model<MODEL, NPL, 32>(ob, um, ptab, pr.phi_min, has_spare, gmask, lig, t1, gm, ll, okm);
#pragma unroll
            for (int i = 0; i < D; ++i) H[i][j] = (okp && okm) ? -(gp[i] - gm[i]) / (2.0 * h) : (i == j ? 1.0 : 0.0);
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < i; ++j) { double s = 0.5 * (H[i][j] + H[j][i]); H[i][j] = s; H[j][i] = s; }
        bool accepted = false;
        for (int tries = 0; tries < 40 && !accepted; ++tries) {
            double Hd[D][D], rhs[D], du[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int j = 0; j < D; ++j) Hd[i][j] = H[i][j];
                Hd[i][i] += lambda * (fabs(H[i][i]) + 1e-8);
                rhs[i] = -g[i];
            }
            if (!chol_solve<D>(Hd, rhs, du)) { lambda *= 10.0; continue; }
            double un[D], gn[D], lpn, dmax = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                un[i] = fmin(fmax(u[i] + du[i], -UMAX), UMAX);
                dmax = fmax(dmax, fabs(un[i] - u[i]));
            }
            bool okn;
            eval_model<MODEL, NPL, 32>(ob, un, ptab, pr.phi_min, has_spare, gmask, lig, lpn, gn, ll, okn);
            if (okn && -(lpn + sumC) <= f + slack) {
                double gmax = 0.0;
#pragma unroll
                for (int i = 0; i < D; ++i) { u[i] = un[i]; g[i] = -gn[i]; gmax = fmax(gmax, fabs(gn[i])); }
                f = -(lpn + sumC);
                lambda = fmax(lambda / 3.0, 1e-12);
                accepted = true;
                if (dmax < 1e-10 || gmax < 1e-9) converged = true;
            } else {
                lambda *= 4.0;
                if (dmax < 1e-13) { converged = true; accepted = true; }
            }
        }
        if (!accepted) break;
        if (converged) { ++it; break; }
    }
    double th[4];
    constrain<MODEL>(u, pr.phi_min, th);
#pragma unroll
    for (int j = 0; j < 4; ++j) out.theta[j] = th[j];
    out.logp = -f;
    out.iters = (uint32_t)it;
    out.converged = converged ? 1u : 0u;
}

#ifndef MDG_MAP_MINBLOCKS
#define MDG_MAP_MINBLOCKS 4
#endif
template <int NPL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, MDG_MAP_MINBLOCKS) map_kernel(const MapLaunch p) {
    __shared__ unsigned int sh_item[WARPS];
    __shared__ double2 sh_prior[2][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    log_table_init();
    prior_table_init<0>(sh_prior[0], p.pr, 0);
    prior_table_init<1>(sh_prior[1], p.pr, 0);
    for (;;) {
        if (lane == 0) sh_item[warp] = atomicAdd(p.work_counter, 1u);
        __syncwarp();
        const unsigned item = sh_item[warp];
        __syncwarp();
        if (item >= 2u * (unsigned)p.n_tax) break;
        const int tax = (int)(item >> 1), model = (int)(item & 1u);
        LaneObs<NPL> ob;
        load_obs<NPL, 32>(ob, p.k + (size_t)tax * 2 * p.P, p.N + (size_t)tax * 2 * p.P, p.P, 0, lane);
        double logC[NPL];
        log_binom_coeff<NPL>(ob, logC);
        const bool has_spare = 2 * p.P < NPL * 32;
        MapRecord rec;
        if (model == 0) map_fit_group<0, NPL>(ob, logC, p.P, p.pr, sh_prior[0], has_spare, lane, rec);
        else map_fit_group<1, NPL>(ob, logC, p.P, p.pr, sh_prior[1], has_spare, lane, rec);
        if (lane == 0) p.rec[item] = rec;
    }
}

}  // namespace mdg
