// Window-aligned split plan of the propagation kernel (include/hgr.h: hgr_window_split_count / _fill).
//
// Why: the propagation gathers one 256-byte embedding row per nonzero.  When the gathered table is larger than L2 (the 320 MB
// user table of the 1.25 M x 0.25 M graph, the 3 GB table of an 8-GPU block), a row whose columns span the whole table misses
// L2 on almost every gather -- unless all rows walk the table together, window by window.  A chunk of this plan ends where its
// row leaves a window of 2^window_shift table rows, once it holds min_seg nonzeros (and at max_seg at the latest: a chunk of a
// dense row lies inside one window, a chunk of a sparse row spans the few windows its min_seg nonzeros need); the work list orders chunks by
// window (graph.py: _windowed_schedule), so a window is read from DRAM once and then served from L2 to every row that
// references it.  The partial rows cost 512 bytes of traffic per chunk, far less than the ~min_seg x 256 bytes they save.
//
// The walk is sequential inside a row (one thread per row, two passes: count, then fill).  It runs once per graph.
#include "hgr_internal.cuh"

namespace hgr {
namespace {

template <typename Emit>
__device__ __forceinline__ int walk_row(const int32_t *__restrict__ indices, int64_t s, int64_t e, int shift, int min_seg, int max_seg, int min_span,
                                        Emit emit) {
    if (s >= e) return 1;
    int n = 1;
    emit(0, s);
    // a row whose columns stay inside min_span windows gathers from a stretch of the table that L2 holds anyway: only the cap cuts it
    const bool windows = (indices[e - 1] >> shift) - (indices[s] >> shift) >= min_span;
    int64_t cs = s;
    int prev = indices[s] >> shift;
    for (int64_t j = s + 1; j < e; ++j) {
        const int w = indices[j] >> shift;
        const int64_t len = j - cs;
        if (len >= max_seg || (windows && w != prev && len >= min_seg)) {
            emit(n, j);
            ++n;
            cs = j;
        }
        prev = w;
    }
    return n;
}

__global__ void __launch_bounds__(256) window_split_count_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, int n_rows,
                                                                 int shift, int min_seg, int max_seg, int min_span, int32_t *__restrict__ n_seg) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    n_seg[r] = walk_row(indices, indptr[r], indptr[r + 1], shift, min_seg, max_seg, min_span, [](int, int64_t) {});
}

__global__ void __launch_bounds__(256) window_split_fill_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                                const int32_t *__restrict__ heavy_rows, const int64_t *__restrict__ heavy_chunk_ptr,
                                                                int n_heavy, int shift, int min_seg, int max_seg, int min_span, int64_t *__restrict__ chunk_start) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_heavy) return;
    const int r = heavy_rows[h];
    const int64_t c0 = heavy_chunk_ptr[h], c1 = heavy_chunk_ptr[h + 1];
    walk_row(indices, indptr[r], indptr[r + 1], shift, min_seg, max_seg, min_span, [&](int k, int64_t j) {
        if (c0 + k < c1) chunk_start[c0 + k] = j;
    });
}

}  // namespace
}  // namespace hgr

using namespace hgr;

extern "C" {

int hgr_window_split_count(const int64_t *indptr, const int32_t *indices, int32_t n_rows, int32_t window_shift, int32_t min_seg,
                           int32_t max_seg, int32_t min_span, int32_t *n_seg, hgr_stream_t stream) {
    HGR_REQUIRE(n_rows >= 0, "window_split_count: negative n_rows");
    HGR_REQUIRE(window_shift >= 0 && window_shift < 31, "window_split_count: window_shift %d out of range", window_shift);
    HGR_REQUIRE(min_seg >= 1 && max_seg >= min_seg, "window_split_count: need 1 <= min_seg <= max_seg (got %d, %d)", min_seg, max_seg);
    if (n_rows == 0) return HGR_OK;
    HGR_REQUIRE(indptr && n_seg, "window_split_count: NULL argument");
    window_split_count_kernel<<<(unsigned)ceil_div((int64_t)n_rows, 256), 256, 0, stream>>>(indptr, indices, n_rows, window_shift, min_seg, max_seg, min_span, n_seg);
    HGR_LAUNCH_OK("window_split_count_kernel");
    return HGR_OK;
}

int hgr_window_split_fill(const int64_t *indptr, const int32_t *indices, const int32_t *heavy_rows, const int64_t *heavy_chunk_ptr,
                          int32_t n_heavy_rows, int32_t window_shift, int32_t min_seg, int32_t max_seg, int32_t min_span, int64_t *chunk_start,
                          hgr_stream_t stream) {
    HGR_REQUIRE(n_heavy_rows >= 0, "window_split_fill: negative n_heavy_rows");
    HGR_REQUIRE(window_shift >= 0 && window_shift < 31, "window_split_fill: window_shift %d out of range", window_shift);
    HGR_REQUIRE(min_seg >= 1 && max_seg >= min_seg, "window_split_fill: need 1 <= min_seg <= max_seg (got %d, %d)", min_seg, max_seg);
    if (n_heavy_rows == 0) return HGR_OK;
    HGR_REQUIRE(indptr && indices && heavy_rows && heavy_chunk_ptr && chunk_start, "window_split_fill: NULL argument");
    window_split_fill_kernel<<<(unsigned)ceil_div((int64_t)n_heavy_rows, 256), 256, 0, stream>>>(indptr, indices, heavy_rows, heavy_chunk_ptr,
                                                                                                  n_heavy_rows, window_shift, min_seg, max_seg, min_span, chunk_start);
    HGR_LAUNCH_OK("window_split_fill_kernel");
    return HGR_OK;
}

}  // extern "C"
