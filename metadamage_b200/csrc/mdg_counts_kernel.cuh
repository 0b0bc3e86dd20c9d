// mdg_counts_kernel.cuh — K1 counts_reduce: the mismatch matrix -> per-row reference sums, error
// rates, signed positions, per-TaxID y_sum_total, cut flags and the dense k(z)/N(z) batch.
// Replaces counts.py:237-256 (add_reference_counts 86-89, add_error_rates 109-114, position
// handling 117-129, fillna 175-176, y_sum_total 179-204, cuts 207-209) and the dense extraction
// of fits.py:398-419; optionally the noise statistic of fits.py:359-376.
//
// HBM-bound (111 algorithmic bytes per row). Design: one CTA per tile of T rows
// plus an L-row lookahead; the tile's SoA column chunks are brought into shared memory with 1-D
// TMA bulk copies (cp.async.bulk -> UBLKCP) signalled on one mbarrier; the CTA owns every TaxID
// whose first row lies in [row0, row0+T), so no carry crosses tiles and the only inter-CTA
// traffic is a decoupled look-back on the count of kept TaxIDs (stable compaction of the dense
// output). Inside the tile one warp handles one TaxID at a time (lane = row, i.e. position),
// reducing with shuffles — the same lane<->position mapping as the fit kernels.
#pragma once
#include "mdg_common.cuh"

namespace mdg {

constexpr int kCountsThreads = 256;
constexpr int kCountsWarps = kCountsThreads / 32;

enum CountsError : int { CE_NONE = 0, CE_SEGMENT_TOO_LONG = 1, CE_OVERFLOW = 2 };

struct CountsLaunch {
    long long n_rows;
    const long long* tax_id;
    const uint32_t* n_align;
    const uint8_t* is_rev;
    const uint8_t* pos0;
    const uint32_t* counts16;
    long long stride;
    int fwd_ref, fwd_obs, rev_ref, rev_obs;
    int P;
    uint32_t min_align;
    unsigned long long min_y;
    // per-row outputs (any may be NULL)
    uint32_t* n_fwd_row;
    uint32_t* n_rev_row;
    float* f_fwd_row;
    float* f_rev_row;
    int8_t* z_row;
    unsigned long long* y_row;
    uint8_t* keep_row;
    // per-TaxID outputs
    long long* out_tax;
    uint32_t* out_nal;
    long long* out_first;
    uint32_t* out_k;
    uint32_t* out_N;
    double* out_noise;
    // tiling
    int T, L;               // owned rows per tile, lookahead rows
    int ncols;              // count columns staged in shared memory
    int col_id[16];         // which of the 16 columns
    int col_slot[16];       // column -> staged slot (or -1)
    int use_tma;            // all column bases 16-byte aligned
    unsigned int* tile_ticket;
    unsigned long long* tile_state;  // decoupled look-back: flag << 62 | value
    long long* n_tax_out;            // device scalar
    int* error_flag;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
    return __shfl_xor_sync(0xffffffffu, v, o);
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_u64(v, o);
    return v;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kCountsThreads) counts_reduce_kernel(const CountsLaunch p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap = p.T + p.L;  // rows staged per tile (multiple of 16)
    // shared layout: [tax: (cap+2) i64][counts: ncols*cap u32][nal: cap u32][seg_start: cap+4 i32]
    //                [rev: cap u8][pos: cap u8][seg_kept: cap u8][dense: warps*2*2P u32] [misc]
    long long* s_tax = reinterpret_cast<long long*>(smem_raw);       // s_tax[1+i] = tax of staged row i; s_tax[0] = row before
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_tax + cap + 2);
    uint32_t* s_nal = s_cnt + (size_t)p.ncols * cap;
    int* s_seg = reinterpret_cast<int*>(s_nal + cap);
    uint8_t* s_rev = reinterpret_cast<uint8_t*>(s_seg + cap + 4);  // cap % 16 == 0 keeps 16-byte alignment
    uint8_t* s_pos = s_rev + cap;
    uint8_t* s_kept = s_pos + cap;
    uint32_t* s_dense = reinterpret_cast<uint32_t*>(s_kept + cap);   // cap is a multiple of 16 -> aligned
    __shared__ uint64_t s_bar;
    __shared__ unsigned int s_tile;
    __shared__ int s_warp_cnt[kCountsWarps];
    __shared__ int s_nseg, s_nowned;
    __shared__ long long s_base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_tile = atomicAdd(p.tile_ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const long long row0 = (long long)tile * p.T;
    const long long remaining = p.n_rows - row0;
    const int nload = (int)(remaining < cap ? remaining : cap);
    const int nown = nload < p.T ? nload : p.T;
    const bool last_rows = (row0 + nload == p.n_rows);

    // ---------------- stage the tile ----------------
    const bool tma = p.use_tma && (nload % 16 == 0);
    if (tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar)), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t total = (uint32_t)nload * (8u + 4u + 1u + 1u + 4u * (uint32_t)p.ncols);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&s_bar)), "r"(total) : "memory");
            tma_load_1d(s_tax + 2, p.tax_id + row0, (uint32_t)nload * 8u, &s_bar);
            tma_load_1d(s_nal, p.n_align + row0, (uint32_t)nload * 4u, &s_bar);
            tma_load_1d(s_rev, p.is_rev + row0, (uint32_t)nload, &s_bar);
            tma_load_1d(s_pos, p.pos0 + row0, (uint32_t)nload, &s_bar);
            for (int c = 0; c < p.ncols; ++c)
                tma_load_1d(s_cnt + (size_t)c * cap, p.counts16 + (long long)p.col_id[c] * p.stride + row0, (uint32_t)nload * 4u, &s_bar);
            s_tax[1] = row0 > 0 ? p.tax_id[row0 - 1] : 0;
        }
        __syncthreads();
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0) : "memory");
        }
    } else {
        for (int i = tid; i < nload; i += kCountsThreads) {
            s_tax[2 + i] = p.tax_id[row0 + i];
            s_nal[i] = p.n_align[row0 + i];
            s_rev[i] = p.is_rev[row0 + i];
            s_pos[i] = p.pos0[row0 + i];
        }
        for (int c = 0; c < p.ncols; ++c) {
            const uint32_t* src = p.counts16 + (long long)p.col_id[c] * p.stride + row0;
            for (int i = tid; i < nload; i += kCountsThreads) s_cnt[(size_t)c * cap + i] = src[i];
        }
        if (tid == 0) s_tax[1] = row0 > 0 ? p.tax_id[row0 - 1] : 0;
        __syncthreads();
    }
    const long long* taxv = s_tax + 2;  // taxv[i], i in [-1, nload)

    // ---------------- segment heads: ordered compaction of head indices ----------------
    // each warp scans a contiguous range of rows; two passes (count, then write)
    const int per_warp = ((nload + kCountsWarps - 1) / kCountsWarps + 31) & ~31;
    const int w_lo = warp * per_warp, w_hi = min(nload, w_lo + per_warp);
    int cnt = 0;
    for (int b = w_lo; b < w_hi; b += 32) {
        const int i = b + lane;
        const bool head = (i < w_hi) && ((row0 + i == 0) || (taxv[i] != taxv[i - 1]));
        cnt += __popc(__ballot_sync(0xffffffffu, head));
    }
    if (lane == 0) s_warp_cnt[warp] = cnt;
    __syncthreads();
    int off = 0;
    for (int w = 0; w < warp; ++w) off += s_warp_cnt[w];
    for (int b = w_lo; b < w_hi; b += 32) {
        const int i = b + lane;
        const bool head = (i < w_hi) && ((row0 + i == 0) || (taxv[i] != taxv[i - 1]));
        const unsigned m = __ballot_sync(0xffffffffu, head);
        if (head) s_seg[off + __popc(m & ((1u << lane) - 1u))] = i;
        off += __popc(m);
    }
    if (tid == 0) {
        int n = 0;
        for (int w = 0; w < kCountsWarps; ++w) n += s_warp_cnt[w];
        s_nseg = n;
        s_seg[n] = nload;  // terminator
    }
    __syncthreads();
    const int nseg = s_nseg;
    // owned segments: heads in [0, nown). heads are sorted, so count them with a strided scan
    if (tid == 0) s_nowned = 0;
    __syncthreads();
    {
        int local = 0;
        for (int s = tid; s < nseg; s += kCountsThreads) local += (s_seg[s] < nown);
        local = __reduce_add_sync(0xffffffffu, local);
        if (lane == 0 && local) atomicAdd(&s_nowned, local);
    }
    __syncthreads();
    const int nowned = s_nowned;
    // the last owned segment must terminate inside the staged rows (or at the end of the data)
    if (tid == 0 && nowned > 0) {
        const bool terminated = (nowned < nseg) || last_rows;
        if (!terminated) atomicMax(p.error_flag, (int)CE_SEGMENT_TOO_LONG);
    }

    const int P = p.P, R = 2 * P;
    const int sf0 = p.col_slot[p.fwd_ref * 4], sr0 = p.col_slot[p.rev_ref * 4];
    const int skf = p.col_slot[p.fwd_ref * 4 + p.fwd_obs], skr = p.col_slot[p.rev_ref * 4 + p.rev_obs];

    // ---------------- pass 1: per-row values, y_sum_total, cut flags ----------------
    for (int s = warp; s < nowned; s += kCountsWarps) {
        const int a = s_seg[s], b = s_seg[s + 1];
        unsigned long long ysum = 0;
        for (int r = a + lane; r < b; r += 32) {
            unsigned long long nf = 0, nr = 0;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                nf += s_cnt[(size_t)(sf0 + o) * cap + r];
                nr += s_cnt[(size_t)(sr0 + o) * cap + r];
            }
            const uint32_t kf = s_cnt[(size_t)skf * cap + r], kr = s_cnt[(size_t)skr * cap + r];
            const int zabs = (int)s_pos[r] + 1;
            const bool rev = s_rev[r] != 0;
            if (nf > 0xFFFFFFFFull || nr > 0xFFFFFFFFull) atomicMax(p.error_flag, (int)CE_OVERFLOW);
            const long long g = row0 + r;
            if (p.n_fwd_row) p.n_fwd_row[g] = (uint32_t)nf;
            if (p.n_rev_row) p.n_rev_row[g] = (uint32_t)nr;
            if (p.f_fwd_row) p.f_fwd_row[g] = nf ? (float)((double)kf / (double)nf) : 0.0f;
            if (p.f_rev_row) p.f_rev_row[g] = nr ? (float)((double)kr / (double)nr) : 0.0f;
            if (p.z_row) p.z_row[g] = (int8_t)(rev ? -zabs : zabs);
            if (zabs <= P) ysum += rev ? kr : kf;
        }
        ysum = warp_sum_u64(ysum);
        bool any_keep = false;
        for (int r = a + lane; r < b; r += 32) {
            const int zabs = (int)s_pos[r] + 1;
            const bool keep = (s_nal[r] >= p.min_align) && (ysum >= p.min_y) && (zabs <= P);
            const long long g = row0 + r;
            if (p.y_row) p.y_row[g] = ysum;
            if (p.keep_row) p.keep_row[g] = keep ? 1 : 0;
            any_keep |= keep;
        }
        any_keep = __any_sync(0xffffffffu, any_keep);
        if (lane == 0) s_kept[s] = any_keep ? 1 : 0;
    }
    __syncthreads();

    // ---------------- stable compaction index: block scan + decoupled look-back ----------------
    if (warp == 0) {
        int running = 0;
        for (int b0 = 0; b0 < nowned; b0 += 32) {
            const int s = b0 + lane;
            const int v = (s < nowned) ? (int)s_kept[s] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) {
            // publish aggregate, look back, publish inclusive prefix
            const unsigned long long FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, VMASK = (1ull << 62) - 1;
            volatile unsigned long long* st = p.tile_state;
            long long excl = 0;
            if (tile == 0) {
                st[0] = FLAG_INC | (unsigned long long)running;
            } else {
                st[tile] = FLAG_AGG | (unsigned long long)running;
                __threadfence();
                long long t = (long long)tile - 1;
                while (t >= 0) {
                    unsigned long long v;
                    do { v = st[t]; } while ((v >> 62) == 0ull);
                    excl += (long long)(v & VMASK);
                    if ((v >> 62) == 2ull) break;
                    --t;
                }
                st[tile] = FLAG_INC | (unsigned long long)(excl + running);
            }
            __threadfence();
            s_base = excl;
            // only the LAST tile publishes the total (an earlier tile's lookahead can also reach the end of the data)
            if (row0 + p.T >= p.n_rows && p.n_tax_out) *p.n_tax_out = excl + running;
        }
    }
    __syncthreads();
    const long long base = s_base;

    // ---------------- pass 2: dense k/N (+ noise) of the kept TaxIDs ----------------
    uint32_t* dk = s_dense + (size_t)warp * 2 * R;
    uint32_t* dN = dk + R;
    // rank of segment s among kept = number of kept segments before it (recomputed per warp chunk)
    int rank_before = 0;  // kept count in segments [0, chunk start)
    for (int b0 = 0; b0 < nowned; b0 += kCountsWarps) {
        const int s = b0 + warp;
        // count kept among [b0, s) cheaply: kCountsWarps is 8
        int my_rank = rank_before;
        for (int q = b0; q < s && q < nowned; ++q) my_rank += s_kept[q];
        int chunk_kept = 0;
        for (int q = b0; q < b0 + kCountsWarps && q < nowned; ++q) chunk_kept += s_kept[q];
        rank_before += chunk_kept;
        if (s >= nowned || !s_kept[s]) continue;
        const long long o = base + my_rank;
        const int a = s_seg[s], b = s_seg[s + 1];
        for (int i = lane; i < 2 * R; i += 32) dk[i] = 0;
        __syncwarp();
        for (int r = a + lane; r < b; r += 32) {
            const int zabs = (int)s_pos[r] + 1;
            if (zabs > P) continue;
            const bool rev = s_rev[r] != 0;
            const int slot = rev ? P + zabs - 1 : zabs - 1;
            const int c0 = rev ? sr0 : sf0;
            uint32_t nn = 0;
#pragma unroll
            for (int oo = 0; oo < 4; ++oo) nn += s_cnt[(size_t)(c0 + oo) * cap + r];
            atomicAdd(&dk[slot], s_cnt[(size_t)(rev ? skr : skf) * cap + r]);
            atomicAdd(&dN[slot], nn);
        }
        __syncwarp();
        if (p.out_k) for (int i = lane; i < R; i += 32) p.out_k[o * R + i] = dk[i];
        if (p.out_N) for (int i = lane; i < R; i += 32) p.out_N[o * R + i] = dN[i];
        if (lane == 0) {
            if (p.out_tax) p.out_tax[o] = taxv[a];
            if (p.out_nal) p.out_nal[o] = s_nal[a];
            if (p.out_first) p.out_first[o] = row0 + a;
        }
        if (p.out_noise) {
            // fits.py:359-376 over the TaxID's rows with |z| <= P: blank CT on forward rows and GA on
            // reverse rows, divide every column by its nan-mean, nan-std over all / forward / reverse
            const int OFF[12] = {1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14};
            double inv_mean[12];
#pragma unroll
            for (int c = 0; c < 12; ++c) {
                double sm = 0.0, cn = 0.0;
                const int slotc = p.col_slot[OFF[c]];
                for (int r = a + lane; r < b; r += 32) {
                    const int zabs = (int)s_pos[r] + 1;
                    const bool rev = s_rev[r] != 0;
                    const bool blank = (zabs > P) || (!rev && c == 5) || (rev && c == 6);
                    if (!blank) { sm += (double)s_cnt[(size_t)slotc * cap + r]; cn += 1.0; }
                }
                sm = warp_sum_f64(sm); cn = warp_sum_f64(cn);
                inv_mean[c] = (cn > 0.0 && sm > 0.0) ? cn / sm : nan("");
            }
            double s1[3] = {0, 0, 0}, n1[3] = {0, 0, 0};
            for (int r = a + lane; r < b; r += 32) {
                const int zabs = (int)s_pos[r] + 1;
                const bool rev = s_rev[r] != 0;
                if (zabs > P) continue;
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    const bool blank = (!rev && c == 5) || (rev && c == 6);
                    const double v = (double)s_cnt[(size_t)p.col_slot[OFF[c]] * cap + r] * inv_mean[c];
                    if (!blank && !isnan(v)) { s1[0] += v; n1[0] += 1.0; s1[rev ? 2 : 1] += v; n1[rev ? 2 : 1] += 1.0; }
                }
            }
            double mean3[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) { s1[q] = warp_sum_f64(s1[q]); n1[q] = warp_sum_f64(n1[q]); mean3[q] = s1[q] / n1[q]; }
            double s2[3] = {0, 0, 0};
            for (int r = a + lane; r < b; r += 32) {
                const int zabs = (int)s_pos[r] + 1;
                const bool rev = s_rev[r] != 0;
                if (zabs > P) continue;
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    const bool blank = (!rev && c == 5) || (rev && c == 6);
                    const double v = (double)s_cnt[(size_t)p.col_slot[OFF[c]] * cap + r] * inv_mean[c];
                    if (!blank && !isnan(v)) {
                        const double d0 = v - mean3[0], d1 = v - mean3[rev ? 2 : 1];
                        s2[0] += d0 * d0; s2[rev ? 2 : 1] += d1 * d1;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) s2[q] = warp_sum_f64(s2[q]);
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 3; ++q) p.out_noise[o * 3 + q] = n1[q] > 0.0 ? sqrt(s2[q] / n1[q]) : nan("");
            }
        }
        __syncwarp();
    }
}

}  // namespace mdg
