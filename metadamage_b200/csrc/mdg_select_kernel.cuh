// mdg_select_kernel.cuh — N3: `--max-fits` on the device.
// fits.py:736-744 extract_top_max_fits: df_counts.groupby("tax_id")["N_alignments"].sum().nlargest(max_fits),
// then df_counts.query("tax_id in @top"): the max_fits TaxIDs with the largest sum of N_alignments over
// their (kept) rows, returned in df_counts order. pandas' nlargest(keep="first") on the groupby result
// (index sorted by tax_id) breaks ties at the cut by the smaller tax_id.
//
//   K8a top_weight_kernel  one warp per kept TaxID: weight = sum of N_alignments over its kept rows
//   K8b top_hist_kernel    MSB-first radix select, one byte per launch, on the 128-bit key
//                          (weight, ~tax_id): every CTA replays the earlier digits' decisions from the
//                          global histograms (16 x 256 counters), so no state is carried but those
//   K8c top_flag_kernel    key >= threshold -> per-CTA counts; counts_scan_kernel; top_emit_kernel
//                          writes the selected indices in ascending (= df_counts) order
// Integer work, bit-exact; HBM traffic is 16 passes over 16 B per TaxID: negligible next to the fits.
#pragma once
#include "mdg_common.cuh"

namespace mdg {

constexpr int kTopThreads = 256;
constexpr int kTopPasses = 16;  // bytes of the composite key

struct TopLaunch {
    long long n_tax;
    long long n_top;
    const unsigned long long* weight;
    const long long* tax_id;
    unsigned int* hist;         // [16][256]
    int* block_cnt;             // [n_blocks]
    long long* block_base;      // [n_blocks]
    long long* out_index;
};

// byte `b` (0 = most significant) of the key: bytes 0-7 weight, 8-15 the tax id mapped so that a
// SMALLER id gives a LARGER byte string (sign bit flipped for signed order, then complemented)
__device__ __forceinline__ unsigned top_key_byte(unsigned long long w, long long tax, int b) {
    const unsigned long long lo = ~((unsigned long long)tax ^ 0x8000000000000000ull);
    const unsigned long long v = b < 8 ? w : lo;
    return (unsigned)(v >> (8 * (7 - (b & 7)))) & 0xffu;
}

// replay the digit decisions of passes [0, upto): fills prefix[] and returns how many elements are
// still to be taken among those that match the whole prefix
__device__ __forceinline__ long long top_replay(const unsigned int* hist, long long n_top, int upto, unsigned char* prefix) {
    long long remaining = n_top;
    for (int p = 0; p < upto; ++p) {
        const unsigned int* h = hist + p * 256;
        int d = 255;
        for (; d > 0; --d) {
            const long long c = h[d];
            if (c >= remaining) break;
            remaining -= c;
        }
        prefix[p] = (unsigned char)d;
    }
    return remaining;
}

__global__ void __launch_bounds__(kTopThreads) top_weight_kernel(long long n_rows, const long long* __restrict__ tax_row,
                                                                  const uint32_t* __restrict__ nal_row, const uint8_t* __restrict__ keep_row,
                                                                  long long n_tax, const long long* __restrict__ first_row,
                                                                  unsigned long long* __restrict__ weight) {
    const long long t = (blockIdx.x * (long long)kTopThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_tax) return;
    const long long r0 = first_row[t];
    const long long tax = tax_row[r0];
    unsigned long long sum = 0;
    for (long long base = r0; base < n_rows; base += 32) {  // rows of a TaxID are contiguous
        const long long r = base + lane;
        const bool mine = r < n_rows && tax_row[r] == tax;
        if (mine && (keep_row == nullptr || keep_row[r])) sum += nal_row[r];
        if (__ballot_sync(0xffffffffu, mine) != 0xffffffffu) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) weight[t] = sum;
}

__global__ void __launch_bounds__(kTopThreads) top_hist_kernel(const TopLaunch p, int pass) {
    __shared__ unsigned char s_prefix[kTopPasses];
    __shared__ unsigned int s_hist[256];
    if (threadIdx.x == 0) top_replay(p.hist, p.n_top, pass, s_prefix);
    s_hist[threadIdx.x] = 0;  // kTopThreads == 256
    __syncthreads();
    for (long long i = blockIdx.x * (long long)kTopThreads + threadIdx.x; i < p.n_tax; i += (long long)gridDim.x * kTopThreads) {
        const unsigned long long w = p.weight[i];
        const long long tax = p.tax_id[i];
        bool match = true;
        for (int b = 0; b < pass; ++b) match = match && (top_key_byte(w, tax, b) == s_prefix[b]);
        if (match) atomicAdd(&s_hist[top_key_byte(w, tax, pass)], 1u);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&p.hist[pass * 256 + threadIdx.x], s_hist[threadIdx.x]);
}

// key >= threshold (the 16 replayed digits); MODE 0: count per CTA, MODE 1: write indices
template <int MODE>
__global__ void __launch_bounds__(kTopThreads) top_emit_kernel(const TopLaunch p) {
    __shared__ unsigned char s_prefix[kTopPasses];
    __shared__ int s_warp[kTopThreads / 32];
    if (threadIdx.x == 0) top_replay(p.hist, p.n_top, kTopPasses, s_prefix);
    __syncthreads();
    const long long i = blockIdx.x * (long long)kTopThreads + threadIdx.x;
    bool take = false;
    if (i < p.n_tax) {
        const unsigned long long w = p.weight[i];
        const long long tax = p.tax_id[i];
        take = true;  // equal keys are taken too (the threshold element itself)
        for (int b = 0; b < kTopPasses; ++b) {
            const unsigned kb = top_key_byte(w, tax, b), tb = s_prefix[b];
            if (kb != tb) { take = kb > tb; break; }
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (MODE == 0) {
        if (threadIdx.x == 0) {
            int c = 0;
            for (int w2 = 0; w2 < kTopThreads / 32; ++w2) c += s_warp[w2];
            p.block_cnt[blockIdx.x] = c;
        }
    } else if (take) {
        long long off = p.block_base[blockIdx.x] + __popc(bal & ((1u << lane) - 1u));
        for (int w2 = 0; w2 < warp; ++w2) off += s_warp[w2];
        p.out_index[off] = i;
    }
}

}  // namespace mdg
