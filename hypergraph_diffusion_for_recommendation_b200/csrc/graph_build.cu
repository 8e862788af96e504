// Device construction of canonical CSR matrices from interaction lists, sm_100a.
//
// Replaces the host path of the reference's data layer:
//   Interaction.__create_sparse_bipartite_adjacency   data/ui_graph.py:70-84   (A = [[0,R],[R^T,0]])
//   Interaction.__create_sparse_interaction_matrix    data/ui_graph.py:95-112  (R and R^T as CSR)
//   Graph.normalize_graph_mat                         data/graph.py:11-25      ((D^-1/2 A) D^-1/2, D^-1 R)
// scipy's semantics are kept bit for bit: duplicate entries are summed, columns are ascending
// inside a row, and each normalised value is produced by the same two fp32 multiplies
// (row_scale * a) * col_scale, with the scale factors looked up in a table the HOST computed with
// np.power (numpy's float32 pow is not correctly rounded).
//
// Pipeline (all deterministic, no float atomics):
//   1. pack every entry into a 64-bit key  row << cb | col
//   2. LSD radix sort of the keys, 8 bits per pass, only over the significant bits; each pass is
//      histogram -> scan -> stable scatter with keys staged through shared memory so the global
//      writes are contiguous runs per digit
//   3. head flags + exclusive scan give every distinct key its CSR slot; run lengths are the
//      duplicate multiplicities; row offsets come from binary searches over the sorted keys
//   4. scaling kernels apply the degree normalisation
#include "hgr_internal.cuh"

namespace hgr {

typedef unsigned long long u64;

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys
constexpr int kMaxBlocks = 1184;                // 8 blocks per SM on 148 SMs

struct Partition {
    int n_blocks;
    int64_t per_block;  // multiple of kRsTile
};

static Partition make_partition(int64_t n) {
    Partition p;
    int64_t tiles = ceil_div(n, kRsTile);
    int64_t nb = tiles < kMaxBlocks ? tiles : kMaxBlocks;
    if (nb < 1) nb = 1;
    p.n_blocks = (int)nb;
    p.per_block = ceil_div(tiles, nb) * kRsTile;
    return p;
}

// ---------------------------------------------------------------------------------- key packing
__global__ void pack_keys_kernel(const int32_t *__restrict__ rows, const int32_t *__restrict__ cols, int64_t n, int cb,
                                 u64 *__restrict__ keys) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        keys[j] = ((u64)(uint32_t)rows[j] << cb) | (u64)(uint32_t)cols[j];
}

__global__ void pack_bipartite_keys_kernel(const int32_t *__restrict__ u, const int32_t *__restrict__ it, int64_t n_edges,
                                           int32_t n_users, int cb, u64 *__restrict__ keys) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_edges; j += (int64_t)gridDim.x * blockDim.x) {
        const u64 a = (u64)(uint32_t)u[j], b = (u64)(uint32_t)(it[j] + n_users);
        keys[j] = (a << cb) | b;
        keys[n_edges + j] = (b << cb) | a;
    }
}

// ---------------------------------------------------------------------------------- radix sort pass
__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const u64 *__restrict__ keys, int64_t n, int shift,
                                                             int64_t per_block, uint32_t *__restrict__ hist, int nb) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t beg = (int64_t)blockIdx.x * per_block;
    const int64_t end = beg + per_block < n ? beg + per_block : n;
    const int lane = threadIdx.x & 31;
    for (int64_t base = beg; base < end; base += kRsThreads) {
        const int64_t j = base + threadIdx.x;
        const bool valid = j < end;
        const unsigned d = valid ? (unsigned)((keys[j] >> shift) & 255) : 0x1000u + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == __ffs(peers) - 1) atomicAdd(&h[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `n` uint32 counts into int64 offsets, single block (n <= 256 * kMaxBlocks)
__global__ void __launch_bounds__(1024) scan_u32_to_i64_kernel(const uint32_t *__restrict__ in, int n, int64_t *__restrict__ out,
                                                               int64_t *__restrict__ total) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int j = base + threadIdx.x;
        const int64_t v = j < n ? (int64_t)in[j] : 0;
        int64_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sum[w] = s;
        __syncthreads();
        if (w == 0) {
            int64_t ws = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sum[lane] = ws;  // inclusive over warps
        }
        __syncthreads();
        const int64_t before = carry + (w > 0 ? warp_sum[w - 1] : 0) + (s - v);
        if (j < n) out[j] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, int64_t n,
                                                                int shift, int64_t per_block, const int64_t *__restrict__ base,
                                                                int nb) {
    __shared__ int64_t run_base[256];
    __shared__ uint32_t warp_cnt[kRsWarps][256];
    __shared__ uint32_t tile_start[256];
    __shared__ uint32_t tile_cnt[256];
    __shared__ uint32_t scan_tmp[256];
    __shared__ u64 skeys[kRsTile];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    run_base[threadIdx.x] = base[(size_t)threadIdx.x * nb + blockIdx.x];
    const int64_t beg = (int64_t)blockIdx.x * per_block;
    const int64_t end = beg + per_block < n ? beg + per_block : n;
    for (int64_t tile = beg; tile < end; tile += kRsTile) {
#pragma unroll
        for (int k = 0; k < kRsWarps; ++k) warp_cnt[k][threadIdx.x] = 0;
        __syncthreads();
        // phase 1: rank every key among the keys of equal digit that precede it in this warp's chunk
        u64 key[kRsItems];
        uint32_t rank[kRsItems];
        const int64_t chunk = tile + (int64_t)w * (32 * kRsItems);
#pragma unroll
        for (int it = 0; it < kRsItems; ++it) {
            const int64_t j = chunk + it * 32 + lane;
            const bool valid = j < end;
            key[it] = valid ? in[j] : 0;
            const unsigned d = valid ? (unsigned)((key[it] >> shift) & 255) : 0x1000u + lane;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            uint32_t prev = 0;
            if (valid) prev = warp_cnt[w][d];
            rank[it] = prev + __popc(peers & lt_mask);
            __syncwarp();
            if (valid && lane == __ffs(peers) - 1) warp_cnt[w][d] = prev + __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // phase 2: per digit, exclusive prefix over warps and the tile-level exclusive scan over digits
        {
            uint32_t run = 0;
#pragma unroll
            for (int k = 0; k < kRsWarps; ++k) {
                const uint32_t c = warp_cnt[k][threadIdx.x];
                warp_cnt[k][threadIdx.x] = run;
                run += c;
            }
            tile_cnt[threadIdx.x] = run;
            scan_tmp[threadIdx.x] = run;
        }
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const uint32_t t = threadIdx.x >= o ? scan_tmp[threadIdx.x - o] : 0;
            __syncthreads();
            scan_tmp[threadIdx.x] += t;
            __syncthreads();
        }
        tile_start[threadIdx.x] = scan_tmp[threadIdx.x] - tile_cnt[threadIdx.x];
        __syncthreads();
        // phase 3: stage the tile in shared memory in (digit, original order)
#pragma unroll
        for (int it = 0; it < kRsItems; ++it) {
            const int64_t j = chunk + it * 32 + lane;
            if (j < end) {
                const unsigned d = (unsigned)((key[it] >> shift) & 255);
                skeys[tile_start[d] + warp_cnt[w][d] + rank[it]] = key[it];
            }
        }
        __syncthreads();
        // phase 4: contiguous runs per digit go out to global memory
        const int tile_n = (int)(end - tile < (int64_t)kRsTile ? end - tile : (int64_t)kRsTile);
        for (int t = threadIdx.x; t < tile_n; t += kRsThreads) {
            const u64 k = skeys[t];
            const unsigned d = (unsigned)((k >> shift) & 255);
            out[run_base[d] + (int64_t)(t - tile_start[d])] = k;
        }
        __syncthreads();
        run_base[threadIdx.x] += tile_cnt[threadIdx.x];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- dedup / CSR assembly
// phase a: per-block count of heads (distinct keys) over a contiguous range
__global__ void __launch_bounds__(kRsThreads) head_count_kernel(const u64 *__restrict__ keys, int64_t n, int64_t per_block,
                                                                uint32_t *__restrict__ block_heads) {
    __shared__ uint32_t warp_tot[kRsWarps];
    const int64_t beg = (int64_t)blockIdx.x * per_block;
    const int64_t end = beg + per_block < n ? beg + per_block : n;
    uint32_t c = 0;
    for (int64_t j = beg + threadIdx.x; j < end; j += kRsThreads) c += (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int k = 0; k < kRsWarps; ++k) s += warp_tot[k];
        block_heads[blockIdx.x] = s;
    }
}

// phase c: every head writes its column and its position in the sorted key list to its CSR slot
__global__ void __launch_bounds__(kRsThreads) head_emit_kernel(const u64 *__restrict__ keys, int64_t n, int64_t per_block,
                                                               const int64_t *__restrict__ block_base, int cb,
                                                               int32_t *__restrict__ indices, int64_t *__restrict__ head_pos) {
    __shared__ uint32_t warp_tot[kRsWarps];
    __shared__ int64_t running;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = block_base[blockIdx.x];
    __syncthreads();
    const int64_t beg = (int64_t)blockIdx.x * per_block;
    const int64_t end = beg + per_block < n ? beg + per_block : n;
    const u64 col_mask = (cb >= 64) ? ~0ull : ((1ull << cb) - 1ull);
    for (int64_t base = beg; base < end; base += kRsThreads) {
        const int64_t j = base + threadIdx.x;
        u64 k = 0;
        bool head = false;
        if (j < end) {
            k = keys[j];
            head = (j == 0) || (k != keys[j - 1]);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, head);
        if (lane == 0) warp_tot[w] = __popc(ballot);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int q = 0; q < kRsWarps; ++q) {
            const uint32_t c = warp_tot[q];
            if (q < w) before += c;
            total += c;
        }
        if (head) {
            const int64_t p = running + before + __popc(ballot & ((1u << lane) - 1u));
            indices[p] = (int32_t)(k & col_mask);
            head_pos[p] = j;
        }
        __syncthreads();
        if (threadIdx.x == 0) running += total;
        __syncthreads();
    }
}

__device__ __forceinline__ int64_t lower_bound_key(const u64 *__restrict__ keys, int64_t n, u64 target) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// row offsets (in CSR slots) and raw entry counts per row (duplicates included = row sum for unit weights)
__global__ void row_offsets_kernel(const u64 *__restrict__ keys, int64_t n, const int64_t *__restrict__ head_pos,
                                   const int64_t *__restrict__ nnz_ptr, int32_t n_rows, int cb, int64_t *__restrict__ indptr,
                                   int32_t *__restrict__ row_entries) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    const int64_t nnz = *nnz_ptr;
    const int64_t j0 = lower_bound_key(keys, n, (u64)r << cb);
    // CSR slot of the first distinct key at or after j0 = number of heads before j0
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (head_pos[mid] < j0) lo = mid + 1;
        else hi = mid;
    }
    indptr[r] = lo;
    if (r < n_rows && row_entries) {
        const int64_t j1 = lower_bound_key(keys, n, (u64)(r + 1) << cb);
        row_entries[r] = (int32_t)(j1 - j0);
    }
}

__global__ void multiplicity_kernel(const int64_t *__restrict__ head_pos, const int64_t *__restrict__ nnz_ptr, int64_t n,
                                    float *__restrict__ values) {
    const int64_t nnz = *nnz_ptr;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t nxt = p + 1 < nnz ? head_pos[p + 1] : n;
        values[p] = (float)(nxt - head_pos[p]);
    }
}

// ---------------------------------------------------------------------------------- normalisation
__global__ void degree_scale_kernel(const int32_t *__restrict__ deg, int64_t n, const float *__restrict__ lut, int32_t lut_len,
                                    float *__restrict__ out, int32_t *__restrict__ overflow) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int32_t d = deg[j];
        if (d >= 0 && d < lut_len) out[j] = lut[d];
        else {
            out[j] = 0.f;
            if (overflow) atomicAdd(overflow, 1);  // integer flag: the host raises
        }
    }
}

// values[p] = (row_scale[r] * values[p]) * col_scale[c] -- scipy's (D A) D, two roundings
__global__ void csr_scale_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, float *__restrict__ values,
                                 int32_t n_rows, const float *__restrict__ row_scale, const float *__restrict__ col_scale) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float rs = row_scale ? row_scale[r] : 1.f;
        const int64_t e = indptr[r + 1];
        for (int64_t p = indptr[r] + lane; p < e; p += 32) {
            float v = values[p];
            if (row_scale) v = __fmul_rn(rs, v);
            if (col_scale) v = __fmul_rn(v, col_scale[indices[p]]);
            values[p] = v;
        }
    }
}

// ---------------------------------------------------------------------------------- host drivers
static int bits_for(uint32_t max_value) {
    int b = 1;
    while (b < 32 && (max_value >> b) != 0) ++b;
    return b;
}

struct BuildWs {
    u64 *keys_a, *keys_b;
    int64_t *head_pos;
    uint32_t *hist;
    int64_t *hist_base;
    uint32_t *block_heads;
    int64_t *block_base;
};

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static size_t build_ws_bytes(int64_t n) {
    size_t s = 0;
    s += 2 * align_up((size_t)n * 8);
    s += align_up((size_t)(n + 1) * 8);
    s += align_up((size_t)256 * kMaxBlocks * 4) + align_up((size_t)256 * kMaxBlocks * 8);
    s += align_up((size_t)kMaxBlocks * 4) + align_up((size_t)kMaxBlocks * 8);
    return s;
}

static BuildWs carve(void *ws, int64_t n) {
    BuildWs b;
    char *p = reinterpret_cast<char *>(ws);
    b.keys_a = reinterpret_cast<u64 *>(p); p += align_up((size_t)n * 8);
    b.keys_b = reinterpret_cast<u64 *>(p); p += align_up((size_t)n * 8);
    b.head_pos = reinterpret_cast<int64_t *>(p); p += align_up((size_t)(n + 1) * 8);
    b.hist = reinterpret_cast<uint32_t *>(p); p += align_up((size_t)256 * kMaxBlocks * 4);
    b.hist_base = reinterpret_cast<int64_t *>(p); p += align_up((size_t)256 * kMaxBlocks * 8);
    b.block_heads = reinterpret_cast<uint32_t *>(p); p += align_up((size_t)kMaxBlocks * 4);
    b.block_base = reinterpret_cast<int64_t *>(p);
    return b;
}

// sorts ws.keys_a (n keys, `key_bits` significant bits); returns the buffer holding the result
static int radix_sort(BuildWs &b, int64_t n, int key_bits, cudaStream_t st, u64 **sorted) {
    const Partition part = make_partition(n);
    u64 *src = b.keys_a, *dst = b.keys_b;
    for (int shift = 0; shift < key_bits; shift += 8) {
        rs_hist_kernel<<<part.n_blocks, kRsThreads, 0, st>>>(src, n, shift, part.per_block, b.hist, part.n_blocks);
        HGR_LAUNCH_OK("rs_hist_kernel");
        scan_u32_to_i64_kernel<<<1, 1024, 0, st>>>(b.hist, 256 * part.n_blocks, b.hist_base, nullptr);
        HGR_LAUNCH_OK("scan_u32_to_i64_kernel");
        rs_scatter_kernel<<<part.n_blocks, kRsThreads, 0, st>>>(src, dst, n, shift, part.per_block, b.hist_base, part.n_blocks);
        HGR_LAUNCH_OK("rs_scatter_kernel");
        u64 *t = src;
        src = dst;
        dst = t;
    }
    *sorted = src;
    return HGR_OK;
}

static int assemble_csr(BuildWs &b, const u64 *keys, int64_t n, int32_t n_rows, int cb, int64_t *indptr, int32_t *indices,
                        float *values, int32_t *row_entries, int64_t *nnz_out, cudaStream_t st) {
    const Partition part = make_partition(n);
    head_count_kernel<<<part.n_blocks, kRsThreads, 0, st>>>(keys, n, part.per_block, b.block_heads);
    HGR_LAUNCH_OK("head_count_kernel");
    scan_u32_to_i64_kernel<<<1, 1024, 0, st>>>(b.block_heads, part.n_blocks, b.block_base, nnz_out);
    HGR_LAUNCH_OK("scan_u32_to_i64_kernel");
    head_emit_kernel<<<part.n_blocks, kRsThreads, 0, st>>>(keys, n, part.per_block, b.block_base, cb, indices, b.head_pos);
    HGR_LAUNCH_OK("head_emit_kernel");
    row_offsets_kernel<<<(unsigned)ceil_div((int64_t)n_rows + 1, 256), 256, 0, st>>>(keys, n, b.head_pos, nnz_out, n_rows, cb, indptr,
                                                                                   row_entries);
    HGR_LAUNCH_OK("row_offsets_kernel");
    multiplicity_kernel<<<kMaxBlocks, 256, 0, st>>>(b.head_pos, nnz_out, n, values);
    HGR_LAUNCH_OK("multiplicity_kernel");
    return HGR_OK;
}

static int build_common(bool bipartite, const int32_t *rows, const int32_t *cols, int64_t n_in, int32_t n_rows, int32_t n_cols,
                        int32_t n_users, int64_t *indptr, int32_t *indices, float *values, int32_t *row_entries, int64_t *nnz_out,
                        void *ws, size_t ws_bytes, cudaStream_t st) {
    HGR_REQUIRE(n_in >= 0 && n_rows >= 0 && n_cols >= 0, "negative size");
    HGR_REQUIRE(indptr && nnz_out, "indptr or nnz_out is NULL");
    const int64_t n = bipartite ? 2 * n_in : n_in;
    if (n == 0) {
        HGR_CUDA_OK(cudaMemsetAsync(indptr, 0, ((size_t)n_rows + 1) * 8, st));
        HGR_CUDA_OK(cudaMemsetAsync(nnz_out, 0, 8, st));
        if (row_entries && n_rows > 0) HGR_CUDA_OK(cudaMemsetAsync(row_entries, 0, (size_t)n_rows * 4, st));
        return HGR_OK;
    }
    HGR_REQUIRE(rows && cols && indices && values, "rows, cols, indices or values is NULL");
    const size_t need = build_ws_bytes(n);
    if (!ws || ws_bytes < need) return set_error(HGR_ERR_WORKSPACE, "CSR build workspace: need %zu bytes, got %zu", need, ws_bytes);
    HGR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "workspace must be 256-byte aligned");
    BuildWs b = carve(ws, n);
    const int cb = bits_for(n_cols > 0 ? (uint32_t)(n_cols - 1) : 0);
    const int rb = bits_for(n_rows > 0 ? (uint32_t)(n_rows - 1) : 0);
    const int grid = (int)(ceil_div(n_in, 256) < kMaxBlocks * 4 ? ceil_div(n_in, 256) : kMaxBlocks * 4);
    if (bipartite) pack_bipartite_keys_kernel<<<grid, 256, 0, st>>>(rows, cols, n_in, n_users, cb, b.keys_a);
    else pack_keys_kernel<<<grid, 256, 0, st>>>(rows, cols, n_in, cb, b.keys_a);
    HGR_LAUNCH_OK("pack_keys_kernel");
    u64 *sorted = nullptr;
    int rc = radix_sort(b, n, cb + rb, st, &sorted);
    if (rc) return rc;
    return assemble_csr(b, sorted, n, n_rows, cb, indptr, indices, values, row_entries, nnz_out, st);
}

}  // namespace hgr

extern "C" {

size_t hgr_build_csr_workspace_bytes(int64_t n_entries) { return hgr::build_ws_bytes(n_entries < 1 ? 1 : n_entries); }

int hgr_coo_to_csr(const int32_t *rows, const int32_t *cols, int64_t n_entries, int32_t n_rows, int32_t n_cols, int64_t *indptr,
                   int32_t *indices, float *values, int32_t *row_entries, int64_t *nnz_out, void *workspace, size_t workspace_bytes,
                   hgr_stream_t stream) {
    return hgr::build_common(false, rows, cols, n_entries, n_rows, n_cols, 0, indptr, indices, values, row_entries, nnz_out, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

int hgr_bipartite_to_csr(const int32_t *users, const int32_t *items, int64_t n_edges, int32_t n_users, int32_t n_items,
                         int64_t *indptr, int32_t *indices, float *values, int32_t *row_entries, int64_t *nnz_out, void *workspace,
                         size_t workspace_bytes, hgr_stream_t stream) {
    HGR_REQUIRE((int64_t)n_users + (int64_t)n_items < (int64_t)0x7fffffff, "too many nodes for int32 ids");
    const int32_t n = n_users + n_items;
    return hgr::build_common(true, users, items, n_edges, n, n, n_users, indptr, indices, values, row_entries, nnz_out, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

int hgr_degree_scale(const int32_t *degree, int64_t n, const float *lut, int32_t lut_len, float *out, int32_t *overflow,
                     hgr_stream_t stream) {
    HGR_REQUIRE(n >= 0 && lut_len > 0, "bad sizes");
    if (n == 0) return HGR_OK;
    HGR_REQUIRE(degree && lut && out, "degree, lut or out is NULL");
    const int grid = (int)(hgr::ceil_div(n, 256) < hgr::kMaxBlocks * 4 ? hgr::ceil_div(n, 256) : hgr::kMaxBlocks * 4);
    hgr::degree_scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(degree, n, lut, lut_len, out, overflow);
    HGR_LAUNCH_OK("degree_scale_kernel");
    return HGR_OK;
}

int hgr_csr_scale(const int64_t *indptr, const int32_t *indices, float *values, int32_t n_rows, const float *row_scale,
                  const float *col_scale, hgr_stream_t stream) {
    HGR_REQUIRE(n_rows >= 0, "n_rows negative");
    if (n_rows == 0) return HGR_OK;
    HGR_REQUIRE(indptr, "indptr is NULL");
    if (!indices || !values) return HGR_OK;  // a matrix without nonzeros has nothing to scale
    const int64_t blocks = hgr::ceil_div((int64_t)n_rows, 8);
    const int grid = (int)(blocks < hgr::kMaxBlocks * 4 ? blocks : hgr::kMaxBlocks * 4);
    hgr::csr_scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(indptr, indices, values, n_rows, row_scale, col_scale);
    HGR_LAUNCH_OK("csr_scale_kernel");
    return HGR_OK;
}

}  // extern "C"
