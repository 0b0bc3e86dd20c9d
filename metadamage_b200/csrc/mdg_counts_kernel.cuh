// mdg_counts_kernel.cuh — K1 counts_reduce: the mismatch matrix -> per-row reference sums, error
// rates, signed positions, per-TaxID y_sum_total, cut flags and the dense k(z)/N(z) batch.
// Replaces counts.py:237-256 (add_reference_counts 86-89, add_error_rates 109-114, position
// handling 117-129, fillna 175-176, y_sum_total 179-204, cuts 207-209) and the dense extraction
// of fits.py:398-419; optionally the noise statistic of fits.py:359-376.
//
// HBM-bound (111 algorithmic bytes per row). Design:
//   * one CTA per tile of T rows plus an L-row lookahead; the tile's SoA column chunks come into
//     shared memory through 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) on one mbarrier;
//   * the CTA owns every TaxID whose first row lies in [row0, row0+T): no carry crosses tiles,
//     the only inter-CTA traffic is ONE atomicAdd per tile that reserves the tile's block of dense
//     output rows; a one-CTA scan and a block-permute kernel then put the blocks into input order
//     (a decoupled look-back was measured to stall 28-42 % of the warp samples: with 448-row tiles
//     the inclusive prefix cannot propagate as fast as tiles retire);
//   * per-row work is vectorised: one thread handles 4 consecutive rows with 128-bit shared
//     loads and writes the row-local outputs (reference sums, error rates, z) of the tile's
//     nominal rows straight from registers with 128-bit stores (tile starts are 16-byte aligned
//     for every column type); y_sum_total and the cut flag follow once the TaxID sums are known;
//   * per-TaxID sums are a short serial loop of one thread per TaxID over staged values (a warp
//     per TaxID with a shuffle reduction was measured 7 % slower); the
//     dense k/N (and the optional noise) of KEPT TaxIDs use one warp per TaxID, lane = row.
// The first version (one warp per TaxID for everything) was instruction-issue bound at 30 warp
// instructions per row (profiles/r01_counts_ncu.md); this layout needs a few.
#pragma once
#include "mdg_common.cuh"

namespace mdg {

constexpr int kCountsThreads = 128;
constexpr int kCountsWarps = kCountsThreads / 32;
// shared bytes per staged row, without the count columns (4 bytes per staged column on top)
constexpr int kCountsBytesPerRow = 8 + 4 + 1 + 1 + 8 + 4 + 1 + 1 + 4 + 8 + 1 + 1;

enum CountsError : int { CE_NONE = 0, CE_SEGMENT_TOO_LONG = 1, CE_OVERFLOW = 2, CE_CAPACITY = 3 };

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
    int T, L;               // owned rows per tile (multiple of 16), lookahead rows; T + L multiple of 16
    int ncols;              // count columns staged in shared memory
    int col_id[16];         // which of the 16 columns
    int col_slot[16];       // column -> staged slot (or -1)
    int use_tma;            // all column bases 16-byte aligned
    int vec_out;            // per-row output bases 16-byte aligned -> 128-bit stores
    unsigned int* tile_ticket;
    unsigned long long* kept_counter;  // device scalar: dense rows reserved so far (unordered, one atomicAdd per tile)
    long long* tile_base;              // [n_tiles] start of the tile's block in the temporary dense arrays
    int* tile_cnt;                     // [n_tiles] kept TaxIDs of the tile
    long long capacity;                // rows available in the per-TaxID arrays
    int* error_flag;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// IEEE double division, kept out of line: error_rate needs it for about one value in 2^26, but
// inlined it was if-converted and executed for every row (ncu: 14.6 % of the kernel's instructions).
__device__ __noinline__ double exact_div(double a, double b) { return a / b; }

// float32(double(k) / double(n)) with 0/0 -> 0 (counts.py:99, 254), bit-exact.
// Fast path: q = k * (1/n) refined once (Markstein): at most 1 ulp(double) from the correctly
// rounded quotient, which changes the float32 result only if q sits within a few ulps of a
// float32 rounding boundary; exactly those (about 1 in 2^26) values take the IEEE division.
__device__ __forceinline__ float error_rate(uint32_t k, uint32_t n) {
    const double a = (double)(k ? k : 1u), b = (double)(n ? n : 1u);
    const double y = rcp_pos(b);
    const double q0 = a * y;
    double q = fma(fma(-b, q0, a), y, q0);
    const unsigned low = (unsigned)__double2loint(q) & 0x1fffffffu;  // the 29 bits float32 drops
    if (low - 0x0ffffffcu <= 8u) q = exact_div(a, b);                // within 4 ulps of a tie: exact path
    return (k && n) ? (float)q : 0.0f;
}

#ifndef MDG_COUNTS_MINBLOCKS
#define MDG_COUNTS_MINBLOCKS 6
#endif

__global__ void __launch_bounds__(kCountsThreads, MDG_COUNTS_MINBLOCKS) counts_reduce_kernel(const CountsLaunch p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap = p.T + p.L;  // rows staged per tile (multiple of 16)
    // ---- shared layout (every array starts 16-byte aligned because cap % 16 == 0) ----
    long long* s_tax = reinterpret_cast<long long*>(smem_raw);  // s_tax[2+i] = tax id of staged row i; s_tax[1] = row before
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_tax + cap + 2);
    uint32_t* s_nal = s_cnt + (size_t)p.ncols * cap;
    uint32_t* o_nf = s_nal + cap;
    uint32_t* o_nr = o_nf + cap;
    uint32_t* o_yc = o_nr + cap;                                // y contribution of the row; later: rank of the segment
    int* s_seg = reinterpret_cast<int*>(o_yc + cap);            // head row of segment s; s_seg[nseg] = nload
    unsigned long long* s_ysum = reinterpret_cast<unsigned long long*>(s_seg + cap + 4);
    uint8_t* s_rev = reinterpret_cast<uint8_t*>(s_ysum + cap);
    uint8_t* s_pos = s_rev + cap;
    int8_t* o_z = reinterpret_cast<int8_t*>(s_pos + cap);
    uint8_t* o_head = reinterpret_cast<uint8_t*>(o_z + cap);
    uint8_t* s_kept = o_head + cap;
    int* s_gseg = reinterpret_cast<int*>(s_kept + cap);              // [cap / 4] segment of each 4-row group's first row
    uint32_t* s_dense = reinterpret_cast<uint32_t*>(s_gseg + cap / 4);  // [warps][2][2P]
    __shared__ uint64_t s_bar;
    __shared__ int s_warp_cnt[kCountsWarps];
    __shared__ int s_nseg, s_nowned, s_kept_total;
    __shared__ long long s_base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_nowned = 0;  // (made visible by the barrier that ends the staging step)
    const unsigned tile = blockIdx.x;  // tiles are independent: no ticket, no ordering between CTAs
    const long long row0 = (long long)tile * p.T;
    const long long remaining = p.n_rows - row0;
    const int nload = (int)(remaining < cap ? remaining : cap);
    const int nown = nload < p.T ? nload : p.T;
    const bool last_rows = (row0 + nload == p.n_rows);
    const int P = p.P, R = 2 * P;

    // ---------------- stage the tile ----------------
    const bool tma = p.use_tma && (nload % 16 == 0);
    if (tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar)), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t total = (uint32_t)nload * (8u + 4u + 1u + 1u + 4u * (uint32_t)p.ncols);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&s_bar)), "r"(total) : "memory");
            tma_load_1d(s_tax + 2, p.tax_id + row0, (uint32_t)nload * 8u, &s_bar);
            for (int c = 0; c < p.ncols; ++c)
                tma_load_1d(s_cnt + (size_t)c * cap, p.counts16 + (long long)p.col_id[c] * p.stride + row0, (uint32_t)nload * 4u, &s_bar);
            tma_load_1d(s_nal, p.n_align + row0, (uint32_t)nload * 4u, &s_bar);
            tma_load_1d(s_rev, p.is_rev + row0, (uint32_t)nload, &s_bar);
            tma_load_1d(s_pos, p.pos0 + row0, (uint32_t)nload, &s_bar);
            s_tax[1] = row0 > 0 ? p.tax_id[row0 - 1] : 0;
            // one thread sleeps on the barrier; the rest of the CTA waits in bar.sync (no issue slots)
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0), "r"(2000000) : "memory");
            }
        }
        __syncthreads();
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

    const int sf0 = p.col_slot[p.fwd_ref * 4], sr0 = p.col_slot[p.rev_ref * 4];
    const int skf = p.col_slot[p.fwd_ref * 4 + p.fwd_obs], skr = p.col_slot[p.rev_ref * 4 + p.rev_obs];

    // ---------------- phase 1: 4 rows per thread; per-row values + ordered list of TaxID heads ----------------
    const int ngroups = (nload + 3) >> 2;
    int seg_base = 0;  // heads found in earlier trips of this loop
    for (int g0 = 0; g0 < ngroups; g0 += kCountsThreads) {
        const int g = g0 + tid;
        const int i0 = g << 2;
        int nheads = 0;
        uint32_t headbits = 0;
        if (g < ngroups) {
            uint32_t nf[4] = {0, 0, 0, 0}, nr[4] = {0, 0, 0, 0};
            uint32_t ovfmask = 0;
            {
                // a uint32 sum of four terms can only wrap if some term has one of its top two bits
                // set: OR everything, and only then (never on real data) redo the sums with carries
                uint4 a[4], b[4];
                uint32_t any = 0;
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    a[o] = *reinterpret_cast<const uint4*>(s_cnt + (size_t)(sf0 + o) * cap + i0);
                    b[o] = *reinterpret_cast<const uint4*>(s_cnt + (size_t)(sr0 + o) * cap + i0);
                    any |= a[o].x | a[o].y | a[o].z | a[o].w | b[o].x | b[o].y | b[o].z | b[o].w;
                }
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    nf[0] += a[o].x; nf[1] += a[o].y; nf[2] += a[o].z; nf[3] += a[o].w;
                    nr[0] += b[o].x; nr[1] += b[o].y; nr[2] += b[o].z; nr[3] += b[o].w;
                }
                if (any >> 30) {
                    unsigned long long wf[4] = {0, 0, 0, 0}, wr[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        wf[0] += a[o].x; wf[1] += a[o].y; wf[2] += a[o].z; wf[3] += a[o].w;
                        wr[0] += b[o].x; wr[1] += b[o].y; wr[2] += b[o].z; wr[3] += b[o].w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) ovfmask |= (uint32_t)(((wf[j] | wr[j]) >> 32) != 0) << j;
                }
            }
            const uint4 kfv = *reinterpret_cast<const uint4*>(s_cnt + (size_t)skf * cap + i0);
            const uint4 krv = *reinterpret_cast<const uint4*>(s_cnt + (size_t)skr * cap + i0);
            const uint32_t kf[4] = {kfv.x, kfv.y, kfv.z, kfv.w}, kr[4] = {krv.x, krv.y, krv.z, krv.w};
            const uchar4 pv = *reinterpret_cast<const uchar4*>(s_pos + i0);
            const uchar4 rv = *reinterpret_cast<const uchar4*>(s_rev + i0);
            const int pos[4] = {pv.x, pv.y, pv.z, pv.w};
            const bool rev[4] = {rv.x != 0, rv.y != 0, rv.z != 0, rv.w != 0};
            long long tx[5];
            tx[0] = taxv[i0 - 1];
            {
                const longlong2 t01 = *reinterpret_cast<const longlong2*>(taxv + i0);
                const longlong2 t23 = *reinterpret_cast<const longlong2*>(taxv + i0 + 2);
                tx[1] = t01.x; tx[2] = t01.y; tx[3] = t23.x; tx[4] = t23.y;
            }
            float ff[4], fr[4];
            uint32_t yc[4], zpack = 0;
            bool ovf = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j;
                const bool valid = i < nload;
                ovf |= valid && ((ovfmask >> j) & 1u);
                const int zabs = pos[j] + 1;
                ff[j] = error_rate(kf[j], nf[j]);
                fr[j] = error_rate(kr[j], nr[j]);
                yc[j] = (valid && zabs <= P) ? (rev[j] ? kr[j] : kf[j]) : 0u;
                const int z = rev[j] ? -zabs : zabs;
                zpack |= ((uint32_t)(uint8_t)(int8_t)z) << (8 * j);
                const bool head = valid && ((row0 + i == 0) || (tx[j + 1] != tx[j]));
                headbits |= (head ? 1u : 0u) << (8 * j);
                nheads += head ? 1 : 0;
            }
            if (ovf) atomicMax(p.error_flag, (int)CE_OVERFLOW);
            *reinterpret_cast<uint4*>(o_nf + i0) = make_uint4(nf[0], nf[1], nf[2], nf[3]);
            *reinterpret_cast<uint4*>(o_nr + i0) = make_uint4(nr[0], nr[1], nr[2], nr[3]);
            *reinterpret_cast<uint4*>(o_yc + i0) = make_uint4(yc[0], yc[1], yc[2], yc[3]);
            *reinterpret_cast<uint32_t*>(o_z + i0) = zpack;
            *reinterpret_cast<uint32_t*>(o_head + i0) = headbits;
            // row-local outputs of the tile's nominal rows [0, nown): every row has exactly one nominal tile
            const long long gi = row0 + i0;
            if (p.vec_out && i0 + 4 <= nown) {
                if (p.n_fwd_row) *reinterpret_cast<uint4*>(p.n_fwd_row + gi) = make_uint4(nf[0], nf[1], nf[2], nf[3]);
                if (p.n_rev_row) *reinterpret_cast<uint4*>(p.n_rev_row + gi) = make_uint4(nr[0], nr[1], nr[2], nr[3]);
                if (p.f_fwd_row) *reinterpret_cast<float4*>(p.f_fwd_row + gi) = make_float4(ff[0], ff[1], ff[2], ff[3]);
                if (p.f_rev_row) *reinterpret_cast<float4*>(p.f_rev_row + gi) = make_float4(fr[0], fr[1], fr[2], fr[3]);
                if (p.z_row) *reinterpret_cast<uint32_t*>(p.z_row + gi) = zpack;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (i0 + j >= nown) continue;
                    if (p.n_fwd_row) p.n_fwd_row[gi + j] = nf[j];
                    if (p.n_rev_row) p.n_rev_row[gi + j] = nr[j];
                    if (p.f_fwd_row) p.f_fwd_row[gi + j] = ff[j];
                    if (p.f_rev_row) p.f_rev_row[gi + j] = fr[j];
                    if (p.z_row) p.z_row[gi + j] = (int8_t)((zpack >> (8 * j)) & 0xffu);
                }
            }
        }
        // ordered compaction of the head rows: warp scan + block offsets
        int incl = nheads;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp_cnt[warp] = incl;
        __syncthreads();
        int off = seg_base + incl - nheads, total = 0;
#pragma unroll
        for (int w = 0; w < kCountsWarps; ++w) { const int c = s_warp_cnt[w]; off += (w < warp) ? c : 0; total += c; }
        // segment (index into s_seg) that this group's first row belongs to; -1 = a TaxID of the previous tile
        if (g < ngroups) s_gseg[g] = off - ((headbits & 1u) ? 0 : 1);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((headbits >> (8 * j)) & 1u) s_seg[off++] = i0 + j;
        seg_base += total;
        __syncthreads();
    }
    if (tid == 0) { s_nseg = seg_base; s_seg[seg_base] = nload; }
    __syncthreads();
    const int nseg = s_nseg;

    // ---------------- phase 2: one thread per owned TaxID: y_sum_total and the cut ----------------
    {
        int local_owned = 0;
        for (int s = tid; s < nseg; s += kCountsThreads) {
            const int a = s_seg[s];
            if (a >= nown) break;  // heads are sorted: the rest belongs to the next tile
            ++local_owned;
            const int b = s_seg[s + 1];
            unsigned long long ysum = 0;
            bool any_row = false;
            for (int r = a; r < b; ++r) {
                ysum += o_yc[r];
                const int z = o_z[r];
                any_row |= (s_nal[r] >= p.min_align) && ((z < 0 ? -z : z) <= P);
            }
            s_ysum[s] = ysum;
            s_kept[s] = (any_row && ysum >= p.min_y) ? 1 : 0;
        }
        local_owned = __reduce_add_sync(0xffffffffu, local_owned);
        if (lane == 0 && local_owned) atomicAdd(&s_nowned, local_owned);
    }
    __syncthreads();
    const int nowned = s_nowned;
    // the last owned TaxID must end inside the staged rows (or at the end of the data)
    if (tid == 0 && nowned > 0 && !((nowned < nseg) || last_rows)) atomicMax(p.error_flag, (int)CE_SEGMENT_TOO_LONG);

    // ---------------- phase 3 (warp 0): ranks of the kept TaxIDs, publish the tile aggregate ----------------
    uint32_t* s_rank = o_yc;  // o_yc is dead after phase 2
    if (warp == 0) {
        int running = 0;
        for (int b0 = 0; b0 < nowned; b0 += 32) {
            const int s = b0 + lane;
            const int v = (s < nowned) ? (int)s_kept[s] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (s < nowned) s_rank[s] = (uint32_t)(running + incl - v);
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) {
            s_kept_total = running;
            const long long base = running ? (long long)atomicAdd(p.kept_counter, (unsigned long long)running) : 0;
            s_base = base;
            p.tile_base[tile] = base;
            p.tile_cnt[tile] = running;
            if (base + running > p.capacity) atomicMax(p.error_flag, (int)CE_CAPACITY);
        }
        __syncwarp();
    }

    // ---------------- phase 4: y_sum_total and the cut flag of the owned rows ----------------
    if (nowned > 0 && (p.y_row || p.keep_row)) {
        const int a0 = s_seg[0], a1 = s_seg[nowned];
        for (int g = tid + (a0 >> 2); g < ((a1 + 3) >> 2); g += kCountsThreads) {
            const int i0 = g << 2;
            // TaxID of the first owned row of this group (recorded by the head compaction of phase 1)
            const int ifirst = i0 < a0 ? a0 : i0;
            int s = i0 < a0 ? 0 : s_gseg[g];
            const uint32_t headbits = *reinterpret_cast<const uint32_t*>(o_head + i0);
            const uint32_t zpack = *reinterpret_cast<const uint32_t*>(o_z + i0);
            const uint4 nal4 = *reinterpret_cast<const uint4*>(s_nal + i0);
            const uint32_t nal[4] = {nal4.x, nal4.y, nal4.z, nal4.w};
            unsigned long long y[4];
            uint32_t keeppack = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j;
                if (i > ifirst && ((headbits >> (8 * j)) & 1u)) ++s;
                const int z = (int)(int8_t)((zpack >> (8 * j)) & 0xffu);
                const unsigned long long ys = s_ysum[s < nowned ? s : nowned - 1];
                y[j] = ys;
                const bool keep = (nal[j] >= p.min_align) && (ys >= p.min_y) && ((z < 0 ? -z : z) <= P);
                keeppack |= (keep ? 1u : 0u) << (8 * j);
            }
            const long long gi = row0 + i0;
            if (p.vec_out && i0 >= a0 && i0 + 4 <= a1) {
                if (p.y_row) {
                    *reinterpret_cast<ulonglong2*>(p.y_row + gi) = make_ulonglong2(y[0], y[1]);
                    *reinterpret_cast<ulonglong2*>(p.y_row + gi + 2) = make_ulonglong2(y[2], y[3]);
                }
                if (p.keep_row) *reinterpret_cast<uint32_t*>(p.keep_row + gi) = keeppack;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + j;
                    if (i < a0 || i >= a1) continue;
                    if (p.y_row) p.y_row[gi + j] = y[j];
                    if (p.keep_row) p.keep_row[gi + j] = (uint8_t)((keeppack >> (8 * j)) & 1u);
                }
            }
        }
    }

    __syncthreads();
    const long long base = s_base;
    if (s_kept_total == 0) return;

    // ---------------- phase 6: dense k/N (+ noise) of the kept TaxIDs, one warp per TaxID ----------------
    uint32_t* dk = s_dense + (size_t)warp * 2 * R;
    uint32_t* dN = dk + R;
    for (int s = warp; s < nowned; s += kCountsWarps) {
        if (!s_kept[s]) continue;
        const long long o = base + (long long)s_rank[s];
        if (o >= p.capacity) continue;
        const int a = s_seg[s], b = s_seg[s + 1];
        for (int i = lane; i < 2 * R; i += 32) dk[i] = 0;
        __syncwarp();
        for (int r = a + lane; r < b; r += 32) {
            const int zabs = (int)s_pos[r] + 1;
            if (zabs > P) continue;
            const bool rev = s_rev[r] != 0;
            const int slot = rev ? P + zabs - 1 : zabs - 1;
            atomicAdd(&dk[slot], s_cnt[(size_t)(rev ? skr : skf) * cap + r]);
            atomicAdd(&dN[slot], rev ? o_nr[r] : o_nf[r]);
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


// ---------------------------------------------------------------------------------------------
// K1b: exclusive scan of the per-tile kept counts (one CTA) -> final block starts and the total
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) counts_scan_kernel(const int* __restrict__ tile_cnt, long long n_tiles,
                                                           long long* __restrict__ final_base, long long* __restrict__ n_tax_out) {
    // warp w scans a contiguous range with coalesced 32-element steps and a running carry; the 32 warp totals are
    // scanned by warp 0; a second coalesced pass adds each warp's offset
    __shared__ long long s_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long per = ((n_tiles + 31) / 32 + 31) / 32 * 32;  // elements per warp, a multiple of 32
    const long long lo = (long long)warp * per, hi = lo + per < n_tiles ? lo + per : n_tiles;
    long long running = 0;
    for (long long i = lo; i < hi; i += 32) {
        const long long t = i + lane;
        const long long v = t < hi ? (long long)tile_cnt[t] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (t < hi) final_base[t] = running + incl - v;
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_tot[warp] = running;
    __syncthreads();
    if (warp == 0) {
        const long long v = s_tot[lane];
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        s_tot[lane] = incl - v;
        if (lane == 31) *n_tax_out = incl;
    }
    __syncthreads();
    const long long off = s_tot[warp];
    if (off != 0)
        for (long long i = lo + lane; i < hi; i += 32) final_base[i] += off;
}

// Two-level form for many tiles (the streaming kernel has one per 128 rows): every CTA scans a chunk of 1024 counts
// (one per thread, coalesced) and records the chunk total; one more tiny launch scans the chunk totals. The final
// position of tile t is chunk_off[t / 1024] + final_base[t].
constexpr int kScanChunk = 1024;

__global__ void __launch_bounds__(kScanChunk) counts_scan_local_kernel(const int* __restrict__ tile_cnt, long long n_tiles,
                                                                       long long* __restrict__ final_base, long long* __restrict__ chunk_tot) {
    __shared__ long long s_w[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long t = (long long)blockIdx.x * kScanChunk + threadIdx.x;
    const long long v = t < n_tiles ? (long long)tile_cnt[t] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const long long w = s_w[lane];
        long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
        s_w[lane] = wi - w;
        if (lane == 31) chunk_tot[blockIdx.x] = wi;
    }
    __syncthreads();
    if (t < n_tiles) final_base[t] = s_w[warp] + incl - v;
}

__global__ void __launch_bounds__(32) counts_scan_chunks_kernel(long long* __restrict__ chunk_tot, long long n_chunks,
                                                                long long* __restrict__ n_tax_out) {
    const int lane = threadIdx.x;
    long long running = 0;
    for (long long i = 0; i < n_chunks; i += 32) {
        const long long c = i + lane;
        const long long v = c < n_chunks ? chunk_tot[c] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (c < n_chunks) chunk_tot[c] = running + incl - v;  // in place: exclusive offsets
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) *n_tax_out = running;
}

// K1c: move every tile's block of per-TaxID rows from its reserved (unordered) place to input order
struct CountsPermute {
    long long n_tiles;
    const long long* tile_base;
    const long long* final_base;
    const long long* chunk_off;  // two-level scan: offset of tile t's chunk (t / kScanChunk), or NULL
    const int* tile_cnt;
    int R;
    const long long* t_tax; long long* out_tax;
    const uint32_t* t_nal; uint32_t* out_nal;
    const long long* t_first; long long* out_first;
    const uint32_t* t_k; uint32_t* out_k;
    const uint32_t* t_N; uint32_t* out_N;
    const double* t_noise; double* out_noise;
};

constexpr int kPermuteWarps = 4;
constexpr int kPermuteTiles = 8;  // consecutive tiles per warp (a divisor of kScanChunk): their blocks are adjacent in the output

// One warp moves the blocks of kPermuteTiles consecutive tiles. Their destinations are contiguous (the scan), so the
// warp sees one output block of `total` per-TaxID rows whose sources are up to eight pieces; the tile bookkeeping is read
// once (lanes 0-7) and handed round by shuffles, so no load depends on another load, and the k / N rows go four at a
// time — all loads of a round are issued before its stores. (Round 2's first version ran one warp per tile: 78 125 tiny
// warps = 19 532 CTAs for the 10M-row input, CTA dispatch alone took most of its 38 us.)
__global__ void __launch_bounds__(kPermuteWarps * 32, 8) counts_permute_kernel(const CountsPermute p) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long t0 = ((long long)blockIdx.x * kPermuteWarps + (threadIdx.x >> 5)) * kPermuteTiles;
    if (t0 >= p.n_tiles) return;
    const long long t = t0 + (lane & (kPermuteTiles - 1));
    int cnt = 0;
    long long src = 0;
    if (t < p.n_tiles) { cnt = __ldg(p.tile_cnt + t); src = __ldg(p.tile_base + t); }
    const long long dst0 = __ldg(p.final_base + t0) + (p.chunk_off ? __ldg(p.chunk_off + t0 / kScanChunk) : 0);
    int start[kPermuteTiles];
    long long srcs[kPermuteTiles];
    int total = 0;
#pragma unroll
    for (int i = 0; i < kPermuteTiles; ++i) {
        srcs[i] = __shfl_sync(FULL, src, i);
        start[i] = total;
        total += __shfl_sync(FULL, cnt, i);
    }
    if (total == 0) return;
    // source row of row j of the warp's output block (starts are non-decreasing: the last tile that starts at or
    // before j owns it; empty tiles share their start with the next one and are overridden)
    auto source_row = [&](int j) -> long long {
        long long s = srcs[0];
        int st = 0;
#pragma unroll
        for (int i = 1; i < kPermuteTiles; ++i)
            if (j >= start[i]) { s = srcs[i]; st = start[i]; }
        return s + (j - st);
    };
    for (int j = lane; j < total; j += 32) {
        const long long s = source_row(j);
        if (p.out_tax) p.out_tax[dst0 + j] = p.t_tax[s];
        if (p.out_nal) p.out_nal[dst0 + j] = p.t_nal[s];
        if (p.out_first) p.out_first[dst0 + j] = p.t_first[s];
        if (p.out_noise) {
#pragma unroll
            for (int c = 0; c < 3; ++c) p.out_noise[(dst0 + j) * 3 + c] = p.t_noise[s * 3 + c];
        }
    }
    const int R = p.R;
    for (int j0 = 0; j0 < total; j0 += 4) {
        long long s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] = j0 + u < total ? source_row(j0 + u) : -1;
        for (int c = lane; c < R; c += 32) {
            uint32_t a[4] = {0u, 0u, 0u, 0u}, b[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (s[u] >= 0) {
                    if (p.out_k) a[u] = p.t_k[s[u] * R + c];
                    if (p.out_N) b[u] = p.t_N[s[u] * R + c];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (s[u] >= 0) {
                    if (p.out_k) p.out_k[(dst0 + j0 + u) * R + c] = a[u];
                    if (p.out_N) p.out_N[(dst0 + j0 + u) * R + c] = b[u];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1d: row order of df_counts (counts.py:167-172 sort_by_alignments): TaxIDs by N_alignments descending, then
// tax_id descending (a sort of n_tax keys, done by the caller), and inside a TaxID the kept rows by
// z = +1..+P, -1..-P (sort key 1/z for z > 0 else z, descending; stable). The row-level part — one
// destination per kept row, O(rows) — is these two kernels: one warp per TaxID of the OUTPUT order.
// ---------------------------------------------------------------------------------------------
struct OrderLaunch {
    long long n_rows;
    const long long* tax_id_row;
    const int8_t* z_row;
    const uint8_t* keep_row;     // NULL: every row is kept
    long long n_tax;
    const long long* first_row;  // [n_tax] first row of every kept TaxID (input order)
    const long long* tax_order;  // [n_tax] index (into first_row) of the TaxID that comes i-th in the output
    int* len_sorted;             // [n_tax] kept rows of the i-th TaxID of the output
    const long long* out_start;  // [n_tax] exclusive scan of len_sorted
    long long* perm;             // [kept rows] source row of every output row
    int* error_flag;
};

constexpr int kOrderWarps = 4;

// rank key of a signed 1-indexed position: +1, +2, ..., then -1, -2, ...
__device__ __forceinline__ int order_zslot(int z) { return z > 0 ? z - 1 : 128 - z - 1; }

template <int PASS>
__global__ void __launch_bounds__(kOrderWarps * 32) counts_order_kernel(const OrderLaunch p) {
    __shared__ short s_key[kOrderWarps][MDG_MAX_SEGMENT_ROWS];  // zslot of the kept rows, -1 for dropped rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * kOrderWarps + warp;
    if (i >= p.n_tax) return;
    const long long r0 = p.first_row[p.tax_order[i]];
    const long long tid = p.tax_id_row[r0];
    // segment length: rows of a TaxID are contiguous
    int L = 0;
    for (int base = 0; base < MDG_MAX_SEGMENT_ROWS; base += 32) {
        const long long r = r0 + base + lane;
        const bool same = r < p.n_rows && p.tax_id_row[r] == tid;
        const unsigned m = __ballot_sync(0xffffffffu, same);
        if (m != 0xffffffffu) { L = base + __ffs(~m) - 1; break; }
        L = base + 32;
    }
    if (L >= MDG_MAX_SEGMENT_ROWS && r0 + L < p.n_rows && p.tax_id_row[r0 + L] == tid) {
        if (lane == 0) atomicMax(p.error_flag, 1);
        return;
    }
    int kept = 0;
    for (int j = lane; j < L; j += 32) {
        const bool k = p.keep_row == nullptr || p.keep_row[r0 + j] != 0;
        s_key[warp][j] = k ? (short)order_zslot((int)p.z_row[r0 + j]) : (short)-1;
        kept += k;
    }
    __syncwarp();
    if (PASS == 0) {
        kept = __reduce_add_sync(0xffffffffu, kept);
        if (lane == 0) p.len_sorted[i] = kept;
    } else {
        const long long base = p.out_start[i];
        for (int j = lane; j < L; j += 32) {
            const int kj = s_key[warp][j];
            if (kj < 0) continue;
            int rank = 0;
            for (int m = 0; m < L; ++m) {
                const int km = s_key[warp][m];
                rank += (km >= 0) && (km < kj || (km == kj && m < j));
            }
            p.perm[base + rank] = r0 + j;
        }
    }
}

}  // namespace mdg
