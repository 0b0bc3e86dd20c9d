// mdg_tsv_kernel.cuh — K0 tsv_parse: the mismatch-matrix text file -> SoA columns on the device.
// The step BEFORE the hot path: replaces the tokenising half of
// dd.read_csv(filename, sep="\t", header=None, names=columns) (counts.py:229-235), which dominates
// the reference's counts wall time. Both layouts are handled: the 22 header-less columns
// counts.py:37-45 expects and the legacy 20 columns (no tax_name / tax_rank) that
// data/input/data_ancient.txt actually has.
//
// Three launches: (1) newline count per 4 KB block, (2) one-CTA exclusive scan, (3) every thread
// records the line starts of its block slice; then one thread per line parses its ~120 bytes.
// Adjacent threads read adjacent lines, so a warp walks ~4 KB of contiguous text.
#pragma once
#include "mdg_common.cuh"

namespace mdg {

constexpr int kTsvBlockBytes = 4096;
constexpr int kTsvThreads = 128;

enum TsvError : int { TE_NONE = 0, TE_FIELDS = 1, TE_NUMBER = 2, TE_RANGE = 3, TE_CAPACITY = 4 };

struct TsvLaunch {
    const char* text;
    long long n_bytes;
    long long first;        // byte offset of the first data line (after an optional header line)
    int n_cols;             // 20 or 22
    long long capacity;     // rows available in the outputs
    long long n_blocks;
    int* block_cnt;         // [n_blocks] newline counts
    long long* block_base;  // [n_blocks] exclusive scan
    long long* line_start;  // [capacity + 1]
    long long* n_lines;     // device scalar
    long long* tax_id;
    uint32_t* n_align;
    uint8_t* is_rev;
    uint8_t* pos0;
    uint32_t* counts16;     // [16][stride]
    long long stride;
    long long* name_span;   // [rows][2] (offset, length) of tax_name, or NULL
    long long* rank_span;   // [rows][2]
    int* error_flag;        // TsvError
    long long* error_line;
};

// a line ends at '\n'; the byte after it starts the next line
__global__ void __launch_bounds__(kTsvThreads) tsv_count_kernel(const TsvLaunch p) {
    __shared__ int s_cnt[kTsvThreads / 32];
    const long long b0 = p.first + (long long)blockIdx.x * kTsvBlockBytes;
    int cnt = 0;
    for (int i = threadIdx.x; i < kTsvBlockBytes; i += kTsvThreads) {
        const long long j = b0 + i;
        if (j < p.n_bytes && p.text[j] == '\n') ++cnt;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kTsvThreads / 32; ++w) t += s_cnt[w];
        p.block_cnt[blockIdx.x] = t;
    }
}

// line k (k >= 1) starts right after the k-th newline; line 0 starts at p.first
__global__ void __launch_bounds__(kTsvThreads) tsv_lines_kernel(const TsvLaunch p) {
    __shared__ int s_warp[kTsvThreads / 32];
    const long long b0 = p.first + (long long)blockIdx.x * kTsvBlockBytes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long base = p.block_base[blockIdx.x] + 1;  // +1: line 0 is implicit
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.capacity >= 0) p.line_start[0] = p.first;
    // ordered: the CTA walks its block in 128-byte steps, one byte per thread
    for (int off = 0; off < kTsvBlockBytes; off += kTsvThreads) {
        const long long j = b0 + off + threadIdx.x;
        const bool nl = j < p.n_bytes && p.text[j] == '\n';
        const unsigned m = __ballot_sync(0xffffffffu, nl);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        long long before = base;
        int total = 0;
        for (int w = 0; w < kTsvThreads / 32; ++w) { const int c = s_warp[w]; if (w < warp) before += c; total += c; }
        if (nl) {
            const long long idx = before + __popc(m & ((1u << lane) - 1u));
            if (idx <= p.capacity) p.line_start[idx] = j + 1;
        }
        base += total;
        __syncthreads();
    }
}

__device__ __forceinline__ bool tsv_parse_uint(const char* t, long long& i, long long end, unsigned long long& v) {
    v = 0;
    int nd = 0;
    while (i < end) {
        const char ch = t[i];
        if (ch < '0' || ch > '9') break;
        v = v * 10ull + (unsigned long long)(ch - '0');
        ++i; ++nd;
        if (nd > 19) return false;
    }
    return nd > 0;
}

__global__ void __launch_bounds__(kTsvThreads) tsv_parse_kernel(const TsvLaunch p, long long n_lines) {
    const long long row = (long long)blockIdx.x * kTsvThreads + threadIdx.x;
    if (row >= n_lines) return;
    const char* t = p.text;
    long long i = p.line_start[row];
    long long end = (row + 1 < n_lines) ? p.line_start[row + 1] - 1 : p.n_bytes;  // excludes the '\n'
    while (end > i && (t[end - 1] == '\n' || t[end - 1] == '\r')) --end;
    int err = TE_NONE;
    auto expect_tab = [&]() { if (i < end && t[i] == '\t') ++i; else err = err ? err : TE_FIELDS; };
    // tax_id (may be negative)
    bool neg = false;
    if (i < end && t[i] == '-') { neg = true; ++i; }
    unsigned long long v;
    if (!tsv_parse_uint(t, i, end, v)) err = TE_NUMBER;
    p.tax_id[row] = neg ? -(long long)v : (long long)v;
    expect_tab();
    if (p.n_cols == 22) {
        for (int f = 0; f < 2; ++f) {
            const long long s0 = i;
            while (i < end && t[i] != '\t') ++i;
            long long* span = f == 0 ? p.name_span : p.rank_span;
            if (span) { span[2 * row] = s0; span[2 * row + 1] = i - s0; }
            expect_tab();
        }
    }
    if (!tsv_parse_uint(t, i, end, v)) err = err ? err : TE_NUMBER;
    if (v > 0xFFFFFFFFull) err = err ? err : TE_RANGE;
    p.n_align[row] = (uint32_t)v;
    expect_tab();
    {   // strand: is_reverse = (field != "5'")  (utils.py:254-255)
        const long long s0 = i;
        while (i < end && t[i] != '\t') ++i;
        p.is_rev[row] = !((i - s0) == 2 && t[s0] == '5' && t[s0 + 1] == '\'');
        expect_tab();
    }
    if (!tsv_parse_uint(t, i, end, v)) err = err ? err : TE_NUMBER;
    if (v > 254ull) err = err ? err : TE_RANGE;  // position + 1 must fit int8 after the sign flip
    p.pos0[row] = (uint8_t)v;
    for (int c = 0; c < 16; ++c) {
        expect_tab();
        if (!tsv_parse_uint(t, i, end, v)) err = err ? err : TE_NUMBER;
        if (v > 0xFFFFFFFFull) err = err ? err : TE_RANGE;  // utils.py:338-339
        p.counts16[(long long)c * p.stride + row] = (uint32_t)v;
    }
    if (i != end) err = err ? err : TE_FIELDS;
    if (err) {
        if (atomicMax(p.error_flag, err) == TE_NONE) *p.error_line = row;
    }
}

}  // namespace mdg
