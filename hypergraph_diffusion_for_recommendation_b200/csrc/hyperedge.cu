// Dense learned-hyperedge propagation of HCCF, sm_100a.
//
// Replaces HGNNLayer.forward (model/graph/HCCF.py:206-211):
//     edge_embeds  = torch.mm(adj.T, embeds)      # [K, D] = H^T E, contraction over ALL n users (or items)
//     hyper_embeds = torch.mm(adj, edge_embeds)   # [n, D] = H T
// with H = dropout(E0 W) [n, K] (K = hyper_dim learned hyperedges, 128 in conf/HCCF.conf) and E [n, D] the layer's
// embeddings.  Both products are "tall and skinny": one dimension is the node count, the others are <= 256.  cuBLAS runs
// the first one as a 64 x 64-tile SIMT sgemm with the whole K = n contraction inside ONE or two thread blocks (0.11 ms on
// a 148-SM device for 0.5 GFLOP); here every SM reduces a slice of the rows into a [K, D] partial in registers and a
// second kernel adds the partials in block order (deterministic, no atomics).  The second product keeps the small matrix
// in shared memory and streams the rows.  The same two kernels serve the backward pass:
//     dT = H^T dY,   dE = H dT,   dH = [dY | E] [T^T ; dT^T]
// Arithmetic is fp32 FMA throughout; results agree with torch.mm to rounding (the summation order differs), tolerance
// 1e-5 relative in the tests.
#include <string.h>

#include "hgr_internal.cuh"

namespace hgr {

constexpr int kHeThreads = 256;
constexpr int kHeRows = 32;  // rows staged per step of the reduce kernel

__device__ __forceinline__ void he_cp_async16(void *smem, const void *gmem, bool valid) {
    // 16-byte asynchronous copy; an invalid source copies nothing and zero-fills the destination
    const unsigned n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(n)
                 : "memory");
}
__device__ __forceinline__ void he_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void he_cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// T_partial[block][K][D] = sum over the block's rows of H[r, :]^T E[r, :].
// Thread (tk, td) of a 16 x 16 layout owns the KT x DT block T[tk * KT ..][td * DT ..]; K = 16 KT, D = 16 DT.
template <int KT, int DT>
__global__ void __launch_bounds__(kHeThreads) tall_skinny_tn_kernel(const float *__restrict__ H, const float *__restrict__ E,
                                                                    int64_t n, int ldh, float *__restrict__ partials) {
    constexpr int K = 16 * KT, D = 16 * DT;
    extern __shared__ __align__(16) float ts_smem[];  // 2 x ([kHeRows][K] rows of H | [kHeRows][D] rows of E)
    constexpr int kTile = kHeRows * (K + D);
    const int td = threadIdx.x % 16, tk = threadIdx.x / 16;
    float acc[KT][DT];
#pragma unroll
    for (int a = 0; a < KT; ++a)
#pragma unroll
        for (int b = 0; b < DT; ++b) acc[a][b] = 0.f;
    // contiguous slice of rows per block: the partial of block b covers rows [b * per, (b + 1) * per)
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r_begin = (int64_t)blockIdx.x * per;
    const int64_t r_end = r_begin + per < n ? r_begin + per : n;
    // the next 32-row tile streams in (cp.async, rows past the slice zero-filled) while the current one is accumulated
    auto prefetch = [&](int64_t r0, float *buf) {
        float *Hs = buf, *Es = buf + kHeRows * K;
        for (int e = threadIdx.x; e < kHeRows * (K / 4); e += kHeThreads) {
            const int r = e / (K / 4), c = e % (K / 4);
            const bool ok = r0 + r < r_end;
            he_cp_async16(Hs + r * K + 4 * c, H + (ok ? r0 + r : 0) * ldh + 4 * c, ok);
        }
        for (int e = threadIdx.x; e < kHeRows * (D / 4); e += kHeThreads) {
            const int r = e / (D / 4), c = e % (D / 4);
            const bool ok = r0 + r < r_end;
            he_cp_async16(Es + r * D + 4 * c, E + (ok ? r0 + r : 0) * D + 4 * c, ok);
        }
        he_cp_async_commit();
    };
    if (r_begin < r_end) prefetch(r_begin, ts_smem);
    int it = 0;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += kHeRows, ++it) {
        const float *Hs = ts_smem + (it & 1) * kTile, *Es = Hs + kHeRows * K;
        if (r0 + kHeRows < r_end) {
            prefetch(r0 + kHeRows, ts_smem + ((it + 1) & 1) * kTile);
            he_cp_async_wait<1>();
        } else {
            he_cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < kHeRows; ++r) {  // zero-filled rows add nothing
            float h[KT], x[DT];
#pragma unroll
            for (int a = 0; a < KT; ++a) h[a] = Hs[r * K + tk * KT + a];
#pragma unroll
            for (int b = 0; b < DT; ++b) x[b] = Es[r * D + td * DT + b];
#pragma unroll
            for (int a = 0; a < KT; ++a)
#pragma unroll
                for (int b = 0; b < DT; ++b) acc[a][b] = fmaf(h[a], x[b], acc[a][b]);
        }
        __syncthreads();  // the buffer just read is the target of the prefetch after next
    }
    float *out = partials + (size_t)blockIdx.x * K * D;
#pragma unroll
    for (int a = 0; a < KT; ++a)
#pragma unroll
        for (int b = 0; b < DT; ++b) out[(tk * KT + a) * D + td * DT + b] = acc[a][b];
}

// T[e] = sum_b partials[b][e], blocks added in order
__global__ void __launch_bounds__(256) partial_sum_kernel(const float *__restrict__ partials, int n_blocks, int n_elems,
                                                          float *__restrict__ T) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    // eight independent chains (blocks b = c mod 8) keep eight loads in flight instead of one dependent load per partial,
    // then a fixed combine: still one order, whatever the launch
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int b = 0;
    for (; b + 8 <= n_blocks; b += 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) s[c] += partials[(size_t)(b + c) * n_elems + e];
    }
    for (int c = 0; b < n_blocks; ++b, ++c) s[c] += partials[(size_t)b * n_elems + e];
    T[e] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// Y[n, N] = [relu]([A1 | A2] B + bias) with B [(K1 + K2), N] resident in shared memory.  A block walks 64-row tiles; the rows
// of the NEXT tile stream into the other half of a double buffer (cp.async) while the current tile is multiplied.
// Thread = 4 output columns x (TR / TY) rows, TY = 256 / (N / 4) row groups; TR = 64 rows per tile, 32 when B is large.
template <int N, int TR>
__global__ void __launch_bounds__(kHeThreads) rows_times_small_kernel(const float *__restrict__ A1, int K1,
                                                                      const float *__restrict__ A2, int K2,
                                                                      const float *__restrict__ B, int64_t n,
                                                                      float *__restrict__ Y, const float *__restrict__ bias,
                                                                      int relu, hgr_gather_t gt) {
    constexpr int TX = N / 4, TY = kHeThreads / TX, RPT = TR / TY;  // rows per thread
    static_assert(RPT >= 1, "tile shorter than the thread layout");
    extern __shared__ __align__(16) float he_smem[];
    const int Kc = K1 + K2;
    float *Bs = he_smem;           // [Kc][N]
    float *As0 = he_smem + Kc * N; // 2 x [TR][Kc]
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    auto prefetch = [&](int64_t r0, float *As) {
        for (int e = threadIdx.x; e < TR * (Kc / 4); e += kHeThreads) {
            const int r = e / (Kc / 4), c = (e % (Kc / 4)) * 4;
            const bool ok = r0 + r < n;
            const int64_t row = ok ? r0 + r : 0;
            const float *src = c < K1 ? A1 + row * K1 + c : A2 + row * K2 + (c - K1);
            he_cp_async16(As + r * Kc + c, src, ok);
        }
        he_cp_async_commit();
    };
    const int64_t step = (int64_t)gridDim.x * TR;
    int64_t r0 = (int64_t)blockIdx.x * TR;
    if (r0 < n) prefetch(r0, As0);
    for (int e = threadIdx.x; e < Kc * (N / 4); e += kHeThreads)
        reinterpret_cast<float4 *>(Bs)[e] = __ldg(reinterpret_cast<const float4 *>(B) + e);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = __ldg(reinterpret_cast<const float4 *>(bias) + tx);
    for (int it = 0; r0 < n; r0 += step, ++it) {
        const float *As = As0 + (it & 1) * TR * Kc;
        if (r0 + step < n) {
            prefetch(r0 + step, As0 + ((it + 1) & 1) * TR * Kc);
            he_cp_async_wait<1>();  // everything but the tile just requested has landed
        } else {
            he_cp_async_wait<0>();
        }
        __syncthreads();
        float4 acc[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < Kc; k += 4) {
            float4 b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4 *>(Bs + (k + j) * N + tx * 4);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const float4 a = *reinterpret_cast<const float4 *>(As + (ty + TY * i) * Kc + k);
                acc[i].x = fmaf(a.x, b[0].x, acc[i].x); acc[i].y = fmaf(a.x, b[0].y, acc[i].y);
                acc[i].z = fmaf(a.x, b[0].z, acc[i].z); acc[i].w = fmaf(a.x, b[0].w, acc[i].w);
                acc[i].x = fmaf(a.y, b[1].x, acc[i].x); acc[i].y = fmaf(a.y, b[1].y, acc[i].y);
                acc[i].z = fmaf(a.y, b[1].z, acc[i].z); acc[i].w = fmaf(a.y, b[1].w, acc[i].w);
                acc[i].x = fmaf(a.z, b[2].x, acc[i].x); acc[i].y = fmaf(a.z, b[2].y, acc[i].y);
                acc[i].z = fmaf(a.z, b[2].z, acc[i].z); acc[i].w = fmaf(a.z, b[2].w, acc[i].w);
                acc[i].x = fmaf(a.w, b[3].x, acc[i].x); acc[i].y = fmaf(a.w, b[3].y, acc[i].y);
                acc[i].z = fmaf(a.w, b[3].z, acc[i].z); acc[i].w = fmaf(a.w, b[3].w, acc[i].w);
            }
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int64_t r = r0 + ty + TY * i;
            float4 o = make_float4(acc[i].x + bv.x, acc[i].y + bv.y, acc[i].z + bv.z, acc[i].w + bv.w);
            if (relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
            if (r < n) {
                reinterpret_cast<float4 *>(Y + r * N)[tx] = o;
                // sharded training: the finished row also goes into every rank's gathered table (the propagation that consumes
                // this layer's output finds it there: no separate exchange)
                if (gt.mc) {
                    st_multicast_f4(reinterpret_cast<float4 *>(gt.mc) + (gt.row_offset + r) * TX + tx, o);
                } else if (gt.n_gather > 0) {
#pragma unroll
                    for (int p = 0; p < HGR_MAX_GATHER; ++p)
                        if (p < gt.n_gather) reinterpret_cast<float4 *>(gt.out[p])[(gt.row_offset + r) * TX + tx] = o;
                }
            }
        }
        __syncthreads();  // the buffer just read is the target of the prefetch after next
    }
}

static int reduce_blocks(int64_t n) {
    int64_t b = ceil_div(n, 4 * kHeRows);  // at least 128 rows per block
    if (b < 1) b = 1;
    if (b > 148 * 2) b = 148 * 2;
    return (int)b;
}

static bool small_dim_ok(int v) { return v == 32 || v == 64 || v == 128; }

template <int KT, int DT>
static int launch_tn_one(int blocks, cudaStream_t st, const float *H, const float *E, int64_t n, int ldh, float *partials) {
    const size_t smem = (size_t)2 * kHeRows * (16 * KT + 16 * DT) * sizeof(float);  // <= 64 KB
    HGR_CUDA_OK(cudaFuncSetAttribute(tall_skinny_tn_kernel<KT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tall_skinny_tn_kernel<KT, DT><<<blocks, kHeThreads, smem, st>>>(H, E, n, ldh, partials);
    return HGR_OK;
}

template <int KT>
static int launch_tn(int D, int blocks, cudaStream_t st, const float *H, const float *E, int64_t n, int ldh, float *partials) {
    switch (D) {
        case 32: return launch_tn_one<KT, 2>(blocks, st, H, E, n, ldh, partials);
        case 64: return launch_tn_one<KT, 4>(blocks, st, H, E, n, ldh, partials);
        default: return launch_tn_one<KT, 8>(blocks, st, H, E, n, ldh, partials);
    }
}

}  // namespace hgr

extern "C" {

size_t hgr_tall_skinny_workspace_bytes(int64_t n, int32_t K, int32_t D) {
    if (n <= 0 || K <= 0 || D <= 0) return 0;
    return (size_t)hgr::reduce_blocks(n) * (size_t)K * (size_t)D * sizeof(float);
}

int hgr_tall_skinny_tn_f32(const float *H, const float *E, int64_t n, int32_t K, int32_t D, float *T, void *workspace,
                           size_t workspace_bytes, hgr_stream_t stream) {
    using namespace hgr;
    HGR_REQUIRE(n >= 0, "n negative");
    HGR_REQUIRE(K == 32 || K == 64 || K == 128 || K == 256, "K = %d unsupported (32, 64, 128 or 256)", K);
    HGR_REQUIRE(small_dim_ok(D), "D = %d unsupported (32, 64 or 128)", D);
    HGR_REQUIRE(T && aligned16(T), "T is NULL or misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        HGR_CUDA_OK(cudaMemsetAsync(T, 0, (size_t)K * D * sizeof(float), st));
        return HGR_OK;
    }
    HGR_REQUIRE(H && E && aligned16(H) && aligned16(E), "H or E is NULL or misaligned");
    const size_t need = hgr_tall_skinny_workspace_bytes(n, K, D);
    if (workspace == nullptr || workspace_bytes < need)
        return set_error(HGR_ERR_WORKSPACE, "tall-skinny workspace: need %zu bytes, got %zu", need, workspace_bytes);
    HGR_REQUIRE(aligned16(workspace), "workspace must be 16-byte aligned");
    const int blocks = reduce_blocks(n);
    float *partials = reinterpret_cast<float *>(workspace);
    // K = 256 runs as two column halves of H (128 accumulators per thread would not fit the register file)
    const int halves = K == 256 ? 2 : 1, Kh = K / halves;
    for (int h = 0; h < halves; ++h) {
        float *part = partials + (size_t)h * blocks * Kh * D;
        int rc;
        switch (Kh) {
            case 32: rc = launch_tn<2>(D, blocks, st, H + h * Kh, E, n, K, part); break;
            case 64: rc = launch_tn<4>(D, blocks, st, H + h * Kh, E, n, K, part); break;
            default: rc = launch_tn<8>(D, blocks, st, H + h * Kh, E, n, K, part); break;
        }
        if (rc) return rc;
        HGR_LAUNCH_OK("tall_skinny_tn_kernel");
        partial_sum_kernel<<<(Kh * D + 255) / 256, 256, 0, st>>>(part, blocks, Kh * D, T + (size_t)h * Kh * D);
        HGR_LAUNCH_OK("partial_sum_kernel");
    }
    return HGR_OK;
}

int hgr_rows_times_small_f32(const float *A1, int32_t K1, const float *A2, int32_t K2, const float *B, int32_t N, int64_t n,
                             float *Y, hgr_stream_t stream) {
    return hgr_rows_times_small_bias_f32(A1, K1, A2, K2, B, N, n, Y, nullptr, 0, stream);
}

int hgr_rows_times_small_bias_f32(const float *A1, int32_t K1, const float *A2, int32_t K2, const float *B, int32_t N, int64_t n,
                                  float *Y, const float *bias, int32_t relu, hgr_stream_t stream) {
    return hgr_rows_times_small_gather_f32(A1, K1, A2, K2, B, N, n, Y, bias, relu, nullptr, stream);
}

int hgr_rows_times_small_gather_f32(const float *A1, int32_t K1, const float *A2, int32_t K2, const float *B, int32_t N, int64_t n,
                                    float *Y, const float *bias, int32_t relu, const hgr_gather_t *gather, hgr_stream_t stream) {
    using namespace hgr;
    hgr_gather_t gt;
    memset(&gt, 0, sizeof(gt));
    if (gather) {
        gt = *gather;
        HGR_REQUIRE(gt.n_gather >= 0 && gt.n_gather <= HGR_MAX_GATHER && gt.row_offset >= 0, "gather: bad n_gather / row_offset");
        for (int j = 0; j < gt.n_gather; ++j) HGR_REQUIRE(gt.out[j] && aligned16(gt.out[j]), "gather: table %d NULL or misaligned", j);
        HGR_REQUIRE(aligned16(gt.mc), "gather: multicast address misaligned");
    }
    HGR_REQUIRE(aligned16(bias), "bias must be 16-byte aligned");
    HGR_REQUIRE(n >= 0, "n negative");
    HGR_REQUIRE(K1 > 0 && K1 % 4 == 0 && K2 >= 0 && K2 % 4 == 0 && K1 + K2 <= 256, "K1 = %d, K2 = %d unsupported", K1, K2);
    HGR_REQUIRE(small_dim_ok(N) || N == 256, "N = %d unsupported (32, 64, 128 or 256)", N);
    if (n == 0) return HGR_OK;
    HGR_REQUIRE(A1 && B && Y && (K2 == 0 || A2), "NULL operand");
    HGR_REQUIRE(aligned16(A1) && aligned16(A2) && aligned16(B) && aligned16(Y), "operands must be 16-byte aligned");
    const int Kc = K1 + K2;
    // B + the double-buffered row tile: 64 rows per tile, 32 when the small matrix leaves less room
    const int TR = ((size_t)Kc * N + (size_t)2 * 64 * Kc) * sizeof(float) <= 200 * 1024 ? 64 : 32;
    const size_t smem = ((size_t)Kc * N + (size_t)2 * TR * Kc) * sizeof(float);
    HGR_REQUIRE(smem <= 220 * 1024, "small matrix [%d, %d] does not fit shared memory", Kc, N);
    int64_t blocks = ceil_div(n, TR);
    if (blocks > 148 * 2) blocks = 148 * 2;
    cudaStream_t st = (cudaStream_t)stream;
#define HGR_RTS(NN, TT)                                                                                                          \
    do {                                                                                                                         \
        HGR_CUDA_OK(cudaFuncSetAttribute(rows_times_small_kernel<NN, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        rows_times_small_kernel<NN, TT><<<(unsigned)blocks, kHeThreads, smem, st>>>(A1, K1, A2, K2, B, n, Y, bias, relu, gt);     \
    } while (0)
    if (TR == 64) {
        switch (N) {
            case 32: HGR_RTS(32, 64); break;
            case 64: HGR_RTS(64, 64); break;
            case 128: HGR_RTS(128, 64); break;
            default: HGR_RTS(256, 64); break;
        }
    } else {
        switch (N) {
            case 32: HGR_RTS(32, 32); break;
            case 64: HGR_RTS(64, 32); break;
            case 128: HGR_RTS(128, 32); break;
            default: HGR_RTS(256, 32); break;
        }
    }
#undef HGR_RTS
    HGR_LAUNCH_OK("rows_times_small_kernel");
    return HGR_OK;
}

}  // extern "C"
