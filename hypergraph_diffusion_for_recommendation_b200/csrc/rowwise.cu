// Row-wise backward of the fused propagation epilogue (LayerNorm o LeakyReLU), sm_100a.
//
// Forward (spmm.cu finish_row): a = leaky_relu(pre); y = (a - mean) * rstd * gamma + beta, the
// lns[k](HGCNConv(...)) pattern of model/graph/HGNN_HD3.py:421,710,714.  Given dy this produces
// dpre and the parameter gradients dgamma = sum_rows dy * xhat, dbeta = sum_rows dy.  The column
// sums are accumulated per block in a fixed row order, written to `partials`, and added up by a
// second kernel in block order: no atomics, bit-reproducible.
#include "hgr_internal.cuh"

namespace hgr {

constexpr int kRwThreads = 256;
constexpr int kMaxPartialBlocks = 1184;  // 8 per SM on a 148-SM B200

template <int LPR>
__global__ void __launch_bounds__(kRwThreads) leaky_ln_bwd_kernel(const float4 *__restrict__ pre,
                                                                   const float4 *__restrict__ dy,
                                                                   const float *__restrict__ gamma, float eps,
                                                                   int use_leaky, float slope, int64_t n_rows,
                                                                   float4 *__restrict__ dpre,
                                                                   float *__restrict__ partials, hgr_gather_t gt) {
    constexpr int GPB = kRwThreads / LPR;
    constexpr int D = LPR * 4;
    constexpr float kInvD = 1.0f / (float)D;
    const int gl = threadIdx.x % LPR;
    const int g = threadIdx.x / LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((threadIdx.x % 32) / LPR * LPR));
    float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
    if (gamma) gm = __ldg(reinterpret_cast<const float4 *>(gamma) + gl);
    float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sb = sg;
    for (int64_t row = (int64_t)blockIdx.x * GPB + g; row < n_rows; row += (int64_t)gridDim.x * GPB) {
        const int64_t off = row * LPR + gl;
        const float4 p = ld_stream_f4(pre + off);
        float4 go = ld_stream_f4(dy + off);
        float4 a = p;
        if (use_leaky) {
            a.x = p.x > 0.f ? p.x : p.x * slope;
            a.y = p.y > 0.f ? p.y : p.y * slope;
            a.z = p.z > 0.f ? p.z : p.z * slope;
            a.w = p.w > 0.f ? p.w : p.w * slope;
        }
        float4 da = go;
        if (gamma) {
            const float mean = group_sum<LPR>((a.x + a.y) + (a.z + a.w), gmask) * kInvD;
            const float cx = a.x - mean, cy = a.y - mean, cz = a.z - mean, cw = a.w - mean;
            const float var = group_sum<LPR>((cx * cx + cy * cy) + (cz * cz + cw * cw), gmask) * kInvD;
            const float rstd = 1.0f / sqrtf(var + eps);
            const float hx = cx * rstd, hy = cy * rstd, hz = cz * rstd, hw = cw * rstd;
            sg.x += go.x * hx;
            sg.y += go.y * hy;
            sg.z += go.z * hz;
            sg.w += go.w * hw;
            sb.x += go.x;
            sb.y += go.y;
            sb.z += go.z;
            sb.w += go.w;
            const float tx = go.x * gm.x, ty = go.y * gm.y, tz = go.z * gm.z, tw = go.w * gm.w;
            const float m1 = group_sum<LPR>((tx + ty) + (tz + tw), gmask) * kInvD;
            const float m2 = group_sum<LPR>((tx * hx + ty * hy) + (tz * hz + tw * hw), gmask) * kInvD;
            da.x = rstd * (tx - m1 - hx * m2);
            da.y = rstd * (ty - m1 - hy * m2);
            da.z = rstd * (tz - m1 - hz * m2);
            da.w = rstd * (tw - m1 - hw * m2);
        }
        if (use_leaky) {
            da.x = p.x > 0.f ? da.x : da.x * slope;
            da.y = p.y > 0.f ? da.y : da.y * slope;
            da.z = p.z > 0.f ? da.z : da.z * slope;
            da.w = p.w > 0.f ? da.w : da.w * slope;
        }
        dpre[off] = da;
        if (gt.mc) {  // through the switch: one multimem store replicated into every rank's gathered table
            st_multicast_f4(reinterpret_cast<float4 *>(gt.mc) + (gt.row_offset + row) * LPR + gl, da);
        } else if (gt.n_gather > 0) {  // fused all-gather: the row into every rank's gathered table (peer-mapped memory)
            const int64_t goff = (gt.row_offset + row) * LPR + gl;
#pragma unroll
            for (int p = 0; p < HGR_MAX_GATHER; ++p)  // constant indices keep gt in param space
                if (p < gt.n_gather) reinterpret_cast<float4 *>(gt.out[p])[goff] = da;
        }
    }
    if (!gamma) return;
    // block reduction over the GPB groups in group order
    __shared__ float4 sh[2][kRwThreads];
    sh[0][threadIdx.x] = sg;
    sh[1][threadIdx.x] = sb;
    __syncthreads();
    if (g == 0) {
        float4 tg = sh[0][gl], tb = sh[1][gl];
        for (int k = 1; k < GPB; ++k) {
            const float4 ug = sh[0][k * LPR + gl], ub = sh[1][k * LPR + gl];
            tg.x += ug.x; tg.y += ug.y; tg.z += ug.z; tg.w += ug.w;
            tb.x += ub.x; tb.y += ub.y; tb.z += ub.z; tb.w += ub.w;
        }
        float4 *out = reinterpret_cast<float4 *>(partials + (size_t)blockIdx.x * 2 * D);
        out[gl] = tg;
        out[LPR + gl] = tb;
    }
}

// One block per parameter column: thread t adds the partials of blocks t, t + 128, ... in order, then a fixed tree over
// the 128 threads -- deterministic, and 100x shorter than one thread walking all partial blocks.
__global__ void __launch_bounds__(128) ln_param_grad_reduce_kernel(const float *__restrict__ partials, int n_blocks, int D,
                                                                    float *__restrict__ dgamma, float *__restrict__ dbeta) {
    __shared__ float sh[128];
    const int k = blockIdx.x;  // 0 .. 2 D - 1
    float s = 0.f;
    for (int b = threadIdx.x; b < n_blocks; b += 128) s += partials[(size_t)b * 2 * D + k];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (k < D) dgamma[k] = sh[0];
        else dbeta[k - D] = sh[0];
    }
}

// y = LayerNorm(x) * gamma + beta, one row per group of D / 4 lanes (the arithmetic of the propagation epilogue).
// Replaces torch's LayerNorm for the MLP input norm of EquivSetConv (model/layers/MLP.py:109-110), which takes
// 2.2 ms for 1.5 M rows of 64 floats; this is a 0.77 GB stream.
template <int LPR>
__global__ void __launch_bounds__(kRwThreads) layer_norm_fwd_kernel(const float4 *__restrict__ x, const float *__restrict__ gamma,
                                                                     const float *__restrict__ beta, float eps, int64_t n_rows,
                                                                     float4 *__restrict__ y) {
    constexpr int GPB = kRwThreads / LPR;
    constexpr float kInvD = 1.0f / (float)(LPR * 4);
    const int gl = threadIdx.x % LPR;
    const int g = threadIdx.x / LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((threadIdx.x % 32) / LPR * LPR));
    const float4 gm = __ldg(reinterpret_cast<const float4 *>(gamma) + gl);
    const float4 bt = __ldg(reinterpret_cast<const float4 *>(beta) + gl);
    for (int64_t row = (int64_t)blockIdx.x * GPB + g; row < n_rows; row += (int64_t)gridDim.x * GPB) {
        const float4 a = ld_stream_f4(x + row * LPR + gl);
        const float mean = group_sum<LPR>((a.x + a.y) + (a.z + a.w), gmask) * kInvD;
        const float cx = a.x - mean, cy = a.y - mean, cz = a.z - mean, cw = a.w - mean;
        const float var = group_sum<LPR>((cx * cx + cy * cy) + (cz * cz + cw * cw), gmask) * kInvD;
        const float rstd = 1.0f / sqrtf(var + eps);
        float4 o;
        o.x = cx * rstd * gm.x + bt.x;
        o.y = cy * rstd * gm.y + bt.y;
        o.z = cz * rstd * gm.z + bt.z;
        o.w = cw * rstd * gm.w + bt.w;
        y[row * LPR + gl] = o;
    }
}

// x[n_rows, D] -> row row_offset + r of every destination table (grid-stride over 128-bit words)
__global__ void __launch_bounds__(kRwThreads) publish_rows_kernel(const float4 *__restrict__ x, int64_t n_words, int64_t word_offset,
                                                                   hgr_gather_t gt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
        const float4 v = ld_stream_f4(x + w);
        if (gt.mc) {
            st_multicast_f4(reinterpret_cast<float4 *>(gt.mc) + word_offset + w, v);
            continue;
        }
#pragma unroll
        for (int p = 0; p < HGR_MAX_GATHER; ++p)
            if (p < gt.n_gather) reinterpret_cast<float4 *>(gt.out[p])[word_offset + w] = v;
    }
}

// out = a + b over 128-bit words, the sum also stored into every rank's gathered table (the "+ res" that ends an encoder layer,
// model/graph/HGNN_HD3.py:419: its result is the input of the next sharded propagation)
__global__ void __launch_bounds__(kRwThreads) add_rows_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, float4 *__restrict__ out,
                                                               int64_t n_words, int64_t word_offset, hgr_gather_t gt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
        const float4 x = ld_stream_f4(a + w), y = ld_stream_f4(b + w);
        const float4 v = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
        out[w] = v;
        if (gt.mc) {
            st_multicast_f4(reinterpret_cast<float4 *>(gt.mc) + word_offset + w, v);
        } else {
#pragma unroll
            for (int p = 0; p < HGR_MAX_GATHER; ++p)
                if (p < gt.n_gather) reinterpret_cast<float4 *>(gt.out[p])[word_offset + w] = v;
        }
    }
}

static int check_gather(const hgr_gather_t *g) {
    HGR_REQUIRE(g != nullptr, "gather is NULL");
    HGR_REQUIRE(g->n_gather >= 0 && g->n_gather <= HGR_MAX_GATHER, "gather: n_gather %d out of range", g->n_gather);
    HGR_REQUIRE(g->row_offset >= 0, "gather: negative row_offset");
    for (int j = 0; j < g->n_gather; ++j) HGR_REQUIRE(g->out[j] && aligned16(g->out[j]), "gather: table %d NULL or misaligned", j);
    HGR_REQUIRE(aligned16(g->mc), "gather: multicast address misaligned");
    return HGR_OK;
}

}  // namespace hgr

extern "C" {

int32_t hgr_ln_bwd_partial_rows(int64_t n_rows) {
    int64_t b = hgr::ceil_div(n_rows, 64);
    if (b < 1) b = 1;
    if (b > hgr::kMaxPartialBlocks) b = hgr::kMaxPartialBlocks;
    return (int32_t)b;
}

int hgr_leaky_ln_bwd_f32(const float *pre, const float *dy, const float *gamma, float ln_eps, int32_t use_leaky,
                         float leaky_slope, int64_t n_rows, int32_t D, float *dpre, float *dgamma, float *dbeta,
                         float *partials, hgr_stream_t stream) {
    hgr_gather_t none;
    none.n_gather = 0;
    none.row_offset = 0;
    none.mc = nullptr;
    return hgr_leaky_ln_bwd_gather_f32(pre, dy, gamma, ln_eps, use_leaky, leaky_slope, n_rows, D, dpre, dgamma, dbeta, partials, &none,
                                       stream);
}

int hgr_publish_rows_f32(const float *x, int64_t n_rows, int32_t D, const hgr_gather_t *gather, hgr_stream_t stream) {
    using namespace hgr;
    HGR_REQUIRE(D > 0 && D % 4 == 0, "D = %d must be a positive multiple of 4", D);
    HGR_REQUIRE(n_rows >= 0, "n_rows negative");
    int rc = check_gather(gather);
    if (rc) return rc;
    if (n_rows == 0 || (gather->n_gather == 0 && !gather->mc)) return HGR_OK;
    HGR_REQUIRE(x && aligned16(x), "x is NULL or misaligned");
    const int64_t n_words = n_rows * (D / 4);
    int64_t blocks = ceil_div(n_words, kRwThreads * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    publish_rows_kernel<<<(unsigned)blocks, kRwThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(x), n_words,
                                                                                 gather->row_offset * (D / 4), *gather);
    HGR_LAUNCH_OK("publish_rows_kernel");
    return HGR_OK;
}

int hgr_add_rows_f32(const float *a, const float *b, int64_t n_rows, int32_t D, float *out, const hgr_gather_t *gather, hgr_stream_t stream) {
    using namespace hgr;
    HGR_REQUIRE(D > 0 && D % 4 == 0 && n_rows >= 0, "bad n_rows / D");
    hgr_gather_t gt;
    gt.n_gather = 0;
    gt.row_offset = 0;
    gt.mc = nullptr;
    if (gather) {
        int rc = check_gather(gather);
        if (rc) return rc;
        gt = *gather;
    }
    if (n_rows == 0) return HGR_OK;
    HGR_REQUIRE(a && b && out && aligned16(a) && aligned16(b) && aligned16(out), "operand NULL or misaligned");
    const int64_t n_words = n_rows * (D / 4);
    int64_t blocks = ceil_div(n_words, kRwThreads * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    add_rows_kernel<<<(unsigned)blocks, kRwThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(a), reinterpret_cast<const float4 *>(b),
                                                                             reinterpret_cast<float4 *>(out), n_words, gt.row_offset * (D / 4), gt);
    HGR_LAUNCH_OK("add_rows_kernel");
    return HGR_OK;
}

int hgr_leaky_ln_bwd_gather_f32(const float *pre, const float *dy, const float *gamma, float ln_eps, int32_t use_leaky,
                                float leaky_slope, int64_t n_rows, int32_t D, float *dpre, float *dgamma, float *dbeta,
                                float *partials, const hgr_gather_t *gather, hgr_stream_t stream) {
    using namespace hgr;
    int grc = check_gather(gather);
    if (grc) return grc;
    const hgr_gather_t gt = *gather;
    HGR_REQUIRE(D == 32 || D == 64 || D == 128, "D = %d unsupported (32, 64 or 128)", D);
    HGR_REQUIRE(n_rows >= 0, "n_rows negative");
    if (n_rows == 0) return HGR_OK;
    HGR_REQUIRE(pre && dy && dpre, "pre, dy or dpre is NULL");
    HGR_REQUIRE(aligned16(pre) && aligned16(dy) && aligned16(dpre) && aligned16(gamma) && aligned16(partials),
                "operands must be 16-byte aligned");
    HGR_REQUIRE(gamma == nullptr || (dgamma && dbeta && partials), "LayerNorm backward needs dgamma, dbeta and partials");
    const int blocks = hgr_ln_bwd_partial_rows(n_rows);
    cudaStream_t st = (cudaStream_t)stream;
    const float4 *p4 = reinterpret_cast<const float4 *>(pre), *g4 = reinterpret_cast<const float4 *>(dy);
    float4 *o4 = reinterpret_cast<float4 *>(dpre);
    switch (D) {
        case 32: leaky_ln_bwd_kernel<8><<<blocks, kRwThreads, 0, st>>>(p4, g4, gamma, ln_eps, use_leaky, leaky_slope, n_rows, o4, partials, gt); break;
        case 64: leaky_ln_bwd_kernel<16><<<blocks, kRwThreads, 0, st>>>(p4, g4, gamma, ln_eps, use_leaky, leaky_slope, n_rows, o4, partials, gt); break;
        default: leaky_ln_bwd_kernel<32><<<blocks, kRwThreads, 0, st>>>(p4, g4, gamma, ln_eps, use_leaky, leaky_slope, n_rows, o4, partials, gt); break;
    }
    HGR_LAUNCH_OK("leaky_ln_bwd_kernel");
    if (gamma) {
        ln_param_grad_reduce_kernel<<<2 * D, 128, 0, st>>>(partials, blocks, D, dgamma, dbeta);
        HGR_LAUNCH_OK("ln_param_grad_reduce_kernel");
    }
    return HGR_OK;
}

int hgr_layer_norm_f32(const float *x, const float *gamma, const float *beta, float ln_eps, int64_t n_rows, int32_t D, float *y,
                       hgr_stream_t stream) {
    using namespace hgr;
    HGR_REQUIRE(D == 32 || D == 64 || D == 128, "D = %d unsupported (32, 64 or 128)", D);
    HGR_REQUIRE(n_rows >= 0, "n_rows negative");
    if (n_rows == 0) return HGR_OK;
    HGR_REQUIRE(x && gamma && beta && y, "NULL argument");
    HGR_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), "operands must be 16-byte aligned");
    int64_t blocks = ceil_div(n_rows, 64);
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    float4 *y4 = reinterpret_cast<float4 *>(y);
    switch (D) {
        case 32: layer_norm_fwd_kernel<8><<<(unsigned)blocks, kRwThreads, 0, st>>>(x4, gamma, beta, ln_eps, n_rows, y4); break;
        case 64: layer_norm_fwd_kernel<16><<<(unsigned)blocks, kRwThreads, 0, st>>>(x4, gamma, beta, ln_eps, n_rows, y4); break;
        default: layer_norm_fwd_kernel<32><<<(unsigned)blocks, kRwThreads, 0, st>>>(x4, gamma, beta, ln_eps, n_rows, y4); break;
    }
    HGR_LAUNCH_OK("layer_norm_fwd_kernel");
    return HGR_OK;
}

}  // extern "C"
