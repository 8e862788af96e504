// Contrastive (SSL) losses fused with their embedding gathers and row normalisation, forward and backward, sm_100a.
//
// Replaces
//   contrastLoss(embeds1, embeds2, nodes, temp)    util/loss_torch.py:103-110, called 2 x n_layers times per batch by
//                                                  HCCF.calcLosses (model/graph/HCCF.py:59-68) and HGNN_HD3.py:345-350
//       a_j = normalize(E1[node_j] + 1e-8),  b_j = normalize(E2[node_j] + 1e-8)        (F.normalize: x / max(||x||, 1e-12))
//       loss = -mean_j log( exp(a_j.b_j / T) / (sum_k exp(a_j.b_k / T) + 1e-8) )
//   InfoNCE(view1, view2, temperature, b_cos)      util/loss_torch.py:32-40, called by SGL.cal_cl_loss (model/graph/SGL.py:167-180)
//       loss = -mean_j log( exp(a_j.b_j / T) / sum_k exp(a_j.b_k / T) + 10e-6 )
// The reference normalises the WHOLE tables (two passes over N x D), materialises the [M, M] logits, and runs
// ~10 kernels; here only the M picked rows are touched and the logits live in registers / shared memory:
//   ssl_prepare_kernel   gather + (+1e-8) + normalise -> A, B [M, D], row norms, the positive dot products
//   ssl_pass_kernel<0>   row sums of exp(logits) (block = 64 rows x all columns, fixed order, no atomics)
//   ssl_finish_kernel    per-row loss terms -> mean (single block, fixed order), per-row backward coefficients
//   ssl_pass_kernel<1/2> backward, logits recomputed tile by tile (flash-style): dA_j = c_j/(M T) (sum_k w_jk b_k - b_j),
//                        dB_k = 1/(M T) (sum_j c_j w_jk a_j - c_k a_k), w_jk = exp(l_jk) / deno_j; then the
//                        normalisation backward and the scatter to the table rows (nodes are unique: plain stores).
#include <math_constants.h>

#include "hgr_internal.cuh"

namespace hgr {

constexpr int SSL_T = 64;         // tile: 64 rows x 64 columns of logits per block step
constexpr int SSL_THREADS = 256;  // 16 x 16 threads, 4 x 4 logits each

// One warp per picked row.
__global__ void __launch_bounds__(256) ssl_prepare_kernel(const float *__restrict__ E1, const float *__restrict__ E2,
                                                          const int64_t *__restrict__ nodes, int64_t M, int64_t n_rows1,
                                                          int64_t n_rows2, int D, float add_eps, int normalize,
                                                          float *__restrict__ A, float *__restrict__ B, float *__restrict__ nrm,
                                                          float *__restrict__ pos, int32_t *__restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= M) return;
    const int64_t r = nodes ? nodes[j] : j;
    const bool ok = r >= 0 && r < n_rows1 && r < n_rows2;
    if (!ok && lane == 0) atomicAdd(bad, 1);
    float xa[4], xb[4], sa = 0.f, sb = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int d = lane + 32 * q;
        xa[q] = (ok && d < D) ? E1[r * D + d] + add_eps : 0.f;
        xb[q] = (ok && d < D) ? E2[r * D + d] + add_eps : 0.f;
        sa += xa[q] * xa[q];
        sb += xb[q] * xb[q];
    }
    sa = group_sum<32>(sa, 0xffffffffu);
    sb = group_sum<32>(sb, 0xffffffffu);
    const float na = normalize ? fmaxf(sqrtf(sa), 1e-12f) : 1.f, nb = normalize ? fmaxf(sqrtf(sb), 1e-12f) : 1.f;
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int d = lane + 32 * q;
        const float ya = xa[q] / na, yb = xb[q] / nb;
        if (d < D) {
            A[j * D + d] = ya;
            B[j * D + d] = yb;
        }
        dot += ya * yb;
    }
    dot = group_sum<32>(dot, 0xffffffffu);
    if (lane == 0) {
        nrm[2 * j] = na;
        nrm[2 * j + 1] = nb;
        pos[j] = dot;
    }
}

// MODE 0: rowsum[j] = sum_k exp(a_j.b_k / T)                     block owns 64 rows j, loops over column tiles
// MODE 1: GA[j, :]  = sum_k c_j exp(a_j.b_k / T) / deno_j * b_k   same ownership
// MODE 2: GB[k, :]  = sum_j c_j exp(a_j.b_k / T) / deno_j * a_j   block owns 64 columns k, loops over row tiles
// X = the owner side's table, Y = the looped side's table (MODE 2: X = B, Y = A).
template <int MODE, int D>
__global__ void __launch_bounds__(SSL_THREADS) ssl_pass_kernel(const float *__restrict__ X, const float *__restrict__ Y, int64_t M,
                                                               float inv_temp, const float *__restrict__ deno,
                                                               const float *__restrict__ coef, float *__restrict__ out) {
    // gridDim.y blocks share an owner tile: block y takes every gridDim.y-th looped tile and writes its own partial
    // result (slice y of `out`); the consumers add the slices in order, so the sums stay deterministic.
    out += (size_t)blockIdx.y * (size_t)M * (MODE == 0 ? 1 : D);
    extern __shared__ float ssl_sm[];
    constexpr int LD = D + 1;
    float *Xs = ssl_sm;               // [64][D + 1] owner rows
    float *Ys = Xs + SSL_T * LD;      // [64][D + 1] looped rows
    float *Gs = Ys + SSL_T * LD;      // [64][65] weights (MODE 1, 2)
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t o0 = (int64_t)blockIdx.x * SSL_T;
    for (int e = threadIdx.x; e < SSL_T * D; e += SSL_THREADS) {
        const int r = e / D, d = e % D;
        Xs[r * LD + d] = o0 + r < M ? X[(o0 + r) * D + d] : 0.f;
    }
    float rowsum[4] = {0.f, 0.f, 0.f, 0.f};
    float acc2[4][D / 16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < D / 16; ++m) acc2[i][m] = 0.f;
    // per-owner-row scale of the weights: MODE 1 c_j / deno_j (rows of this block)
    float oscale[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t j = o0 + ty + 16 * i;
            oscale[i] = j < M ? coef[j] / deno[j] : 0.f;
        }
    }
    for (int64_t l0 = (int64_t)blockIdx.y * SSL_T; l0 < M; l0 += (int64_t)gridDim.y * SSL_T) {
        __syncthreads();
        for (int e = threadIdx.x; e < SSL_T * D; e += SSL_THREADS) {
            const int r = e / D, d = e % D;
            Ys[r * LD + d] = l0 + r < M ? Y[(l0 + r) * D + d] : 0.f;
        }
        __syncthreads();
        // logits of owner rows {ty + 16 i} against looped rows {tx + 16 jj}
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
#pragma unroll 8
        for (int d = 0; d < D; ++d) {
            float xv[4], yv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = Xs[(ty + 16 * i) * LD + d];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) yv[jj] = Ys[(tx + 16 * jj) * LD + d];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(xv[i], yv[jj], acc[i][jj]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int64_t l = l0 + tx + 16 * jj;
                float w = l < M ? expf(acc[i][jj] * inv_temp) : 0.f;
                if (MODE == 0) rowsum[i] += w;
                if (MODE == 1) Gs[(ty + 16 * i) * 65 + tx + 16 * jj] = w * oscale[i];
                if (MODE == 2) Gs[(ty + 16 * i) * 65 + tx + 16 * jj] = l < M ? w * coef[l] / deno[l] : 0.f;  // looped side = rows j
            }
        if (MODE != 0) {
            __syncthreads();
            // acc2[owner row][d] += sum_c G[owner row][c] * Ys[c][d]
#pragma unroll 4
            for (int c = 0; c < SSL_T; ++c) {
                float gv[4], yv[D / 16];
#pragma unroll
                for (int i = 0; i < 4; ++i) gv[i] = Gs[(ty + 16 * i) * 65 + c];
#pragma unroll
                for (int m = 0; m < D / 16; ++m) yv[m] = Ys[c * LD + tx + 16 * m];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int m = 0; m < D / 16; ++m) acc2[i][m] = fmaf(gv[i], yv[m], acc2[i][m]);
            }
        }
    }
    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float s = group_sum<16>(rowsum[i], 0xffffffffu);  // the 16 tx lanes of a half warp share the row
            const int64_t j = o0 + ty + 16 * i;
            if (tx == 0 && j < M) out[j] = s;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t j = o0 + ty + 16 * i;
            if (j < M)
#pragma unroll
                for (int m = 0; m < D / 16; ++m) out[j * D + tx + 16 * m] = acc2[i][m];
        }
    }
}

// kind 0 (contrastLoss): deno = rowsum + 1e-8, term = log(deno) - pos / T, c = 1
// kind 1 (InfoNCE):      r = exp(pos / T) / rowsum, term = -log(r + 10e-6), c = r / (r + 10e-6)
__global__ void __launch_bounds__(1024) ssl_finish_kernel(const float *__restrict__ rowsum, int n_split, const float *__restrict__ pos,
                                                          int64_t M, float inv_temp, int kind, float *__restrict__ deno,
                                                          float *__restrict__ coef, float *__restrict__ loss) {
    __shared__ double sh[1024];
    double t = 0.0;
    for (int64_t j = threadIdx.x; j < M; j += 1024) {
        float rs = rowsum[j];
        for (int s = 1; s < n_split; ++s) rs += rowsum[(int64_t)s * M + j];
        const float p = pos[j] * inv_temp;
        if (kind == 0) {
            const float dn = rs + 1e-8f;
            deno[j] = dn;
            coef[j] = 1.f;
            t += (double)(logf(dn) - p);
        } else {
            const float r = expf(p) / rs;
            deno[j] = rs;
            coef[j] = r / (r + 10e-6f);
            t += (double)(-logf(r + 10e-6f));
        }
    }
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = M > 0 ? (float)(sh[0] / (double)M) : 0.f;
}

// dY_j = scale * (G_j - c_j * other_j)  with scale = grad / (M T); then the normalisation backward
// dx = (dY - y (dY . y)) / ||x|| and the store to the table row (or row j when there is no index).
__global__ void __launch_bounds__(256) ssl_scatter_kernel(const float *__restrict__ G, int n_split, const float *__restrict__ self_n,
                                                          const float *__restrict__ other_n, const float *__restrict__ coef,
                                                          const float *__restrict__ nrm, int which, const int64_t *__restrict__ nodes,
                                                          int64_t M, int64_t n_rows, int D, float inv_temp, int normalize,
                                                          const float *__restrict__ grad_out, float *__restrict__ dE) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= M) return;
    const int64_t r = nodes ? nodes[j] : j;
    if (r < 0 || r >= n_rows) return;
    const float scale = grad_out[0] * inv_temp / (float)M;
    const float c = coef[j];
    float g[4], y[4], gy = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int d = lane + 32 * q;
        y[q] = d < D ? self_n[j * D + d] : 0.f;
        float gs = 0.f;
        if (d < D) {
            gs = G[j * D + d];
            for (int s = 1; s < n_split; ++s) gs += G[((int64_t)s * M + j) * D + d];
        }
        g[q] = d < D ? scale * (gs - c * other_n[j * D + d]) : 0.f;
        gy += g[q] * y[q];
    }
    gy = group_sum<32>(gy, 0xffffffffu);
    const float n = nrm[2 * j + which];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int d = lane + 32 * q;
        if (d < D) dE[r * D + d] = normalize ? (g[q] - y[q] * gy) / n : g[q];
    }
}

constexpr int SSL_MAX_SPLIT = 8;

// slices of the looped dimension: enough blocks for two waves of 148 SMs
static int ssl_splits(int64_t M) {
    const int64_t owners = ceil_div(M > 0 ? M : 1, SSL_T);
    int64_t s = ceil_div(296, owners);
    if (s > SSL_MAX_SPLIT) s = SSL_MAX_SPLIT;
    if (s > owners) s = owners;  // never more slices than looped tiles
    return (int)(s < 1 ? 1 : s);
}

struct SslWs {
    float *A, *B, *nrm, *pos, *rowsum, *deno, *coef, *G;
    size_t total;
};

static SslWs ssl_layout(void *base, int64_t M, int D) {
    SslWs w;
    const size_t S = (size_t)ssl_splits(M);
    size_t o = 0;
    auto take = [&](size_t n) {
        float *p = base ? reinterpret_cast<float *>(static_cast<uint8_t *>(base) + o) : nullptr;
        o += (n * sizeof(float) + 255) / 256 * 256;
        return p;
    };
    const size_t m = (size_t)(M > 0 ? M : 1);
    w.A = take(m * D);
    w.B = take(m * D);
    w.nrm = take(2 * m);
    w.pos = take(m);
    w.rowsum = take(m * S);
    w.deno = take(m);
    w.coef = take(m);
    w.G = take(m * D * S);
    w.total = o;
    return w;
}

template <int MODE>
static int launch_pass(int D, const float *X, const float *Y, int64_t M, float inv_temp, const float *deno, const float *coef,
                       float *out, cudaStream_t st) {
    const dim3 grid((unsigned)ceil_div(M, SSL_T), (unsigned)ssl_splits(M));
    if (grid.x == 0) return HGR_OK;
#define HGR_SSL_LAUNCH(DD)                                                                                                 \
    {                                                                                                                      \
        const size_t smem = (size_t)(2 * SSL_T * (DD + 1) + SSL_T * 65) * sizeof(float);                                   \
        HGR_CUDA_OK(cudaFuncSetAttribute(ssl_pass_kernel<MODE, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ssl_pass_kernel<MODE, DD><<<grid, SSL_THREADS, smem, st>>>(X, Y, M, inv_temp, deno, coef, out);                    \
    }
    if (D == 32) HGR_SSL_LAUNCH(32)
    else if (D == 64) HGR_SSL_LAUNCH(64)
    else HGR_SSL_LAUNCH(128)
#undef HGR_SSL_LAUNCH
    HGR_LAUNCH_OK("ssl_pass_kernel");
    return HGR_OK;
}

}  // namespace hgr

extern "C" {

size_t hgr_ssl_workspace_bytes(int64_t M, int32_t D) { return hgr::ssl_layout(nullptr, M, D).total; }

int hgr_ssl_loss_fwd_f32(const float *E1, const float *E2, int64_t n_rows1, int64_t n_rows2, int32_t D, const int64_t *nodes,
                         int64_t M, float temp, int32_t kind, int32_t normalize, float *loss, void *saved, size_t saved_bytes,
                         int32_t *bad_index_count, hgr_stream_t stream) {
    using namespace hgr;
    cudaStream_t st = (cudaStream_t)stream;
    HGR_REQUIRE(D == 32 || D == 64 || D == 128, "D = %d unsupported (32, 64 or 128)", D);
    HGR_REQUIRE(kind == 0 || kind == 1, "kind %d unknown (0 contrastLoss, 1 InfoNCE)", kind);
    HGR_REQUIRE(temp > 0.f, "temperature must be positive");
    HGR_REQUIRE(M >= 0 && E1 && E2 && loss && bad_index_count, "NULL argument or negative M");
    const SslWs w = ssl_layout(saved, M, D);
    if (saved == nullptr || saved_bytes < w.total)
        return set_error(HGR_ERR_WORKSPACE, "ssl workspace: need %zu bytes, got %zu", w.total, saved_bytes);
    HGR_REQUIRE((reinterpret_cast<uintptr_t>(saved) & 255u) == 0, "workspace must be 256-byte aligned");
    const float inv_temp = 1.0f / temp;
    if (M > 0) {
        ssl_prepare_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(E1, E2, nodes, M, n_rows1, n_rows2, D, kind == 0 ? 1e-8f : 0.f,
                                                                    normalize, w.A, w.B, w.nrm, w.pos, bad_index_count);
        HGR_LAUNCH_OK("ssl_prepare_kernel");
        int rc = launch_pass<0>(D, w.A, w.B, M, inv_temp, nullptr, nullptr, w.rowsum, st);
        if (rc) return rc;
    }
    ssl_finish_kernel<<<1, 1024, 0, st>>>(w.rowsum, ssl_splits(M), w.pos, M, inv_temp, kind, w.deno, w.coef, loss);
    HGR_LAUNCH_OK("ssl_finish_kernel");
    return HGR_OK;
}

int hgr_ssl_loss_bwd_f32(int64_t n_rows1, int64_t n_rows2, int32_t D, const int64_t *nodes, int64_t M, float temp, int32_t normalize,
                         void *saved, size_t saved_bytes, const float *grad_out, float *dE1, float *dE2, hgr_stream_t stream) {
    using namespace hgr;
    cudaStream_t st = (cudaStream_t)stream;
    HGR_REQUIRE(D == 32 || D == 64 || D == 128, "D = %d unsupported (32, 64 or 128)", D);
    HGR_REQUIRE(M >= 0 && grad_out && saved, "NULL argument or negative M");
    const SslWs w = ssl_layout(saved, M, D);
    HGR_REQUIRE(saved_bytes >= w.total, "ssl workspace too small");
    if (M == 0) return HGR_OK;
    const float inv_temp = 1.0f / temp;
    if (dE1) {
        int rc = launch_pass<1>(D, w.A, w.B, M, inv_temp, w.deno, w.coef, w.G, st);
        if (rc) return rc;
        ssl_scatter_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(w.G, ssl_splits(M), w.A, w.B, w.coef, w.nrm, 0, nodes, M, n_rows1, D, inv_temp,
                                                                    normalize, grad_out, dE1);
        HGR_LAUNCH_OK("ssl_scatter_kernel");
    }
    if (dE2) {
        int rc = launch_pass<2>(D, w.B, w.A, M, inv_temp, w.deno, w.coef, w.G, st);
        if (rc) return rc;
        ssl_scatter_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(w.G, ssl_splits(M), w.B, w.A, w.coef, w.nrm, 1, nodes, M, n_rows2, D, inv_temp,
                                                                    normalize, grad_out, dE2);
        HGR_LAUNCH_OK("ssl_scatter_kernel");
    }
    return HGR_OK;
}

}  // extern "C"
