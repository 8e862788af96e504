// BPR ranking loss + L2 regulariser fused with the embedding gathers, forward and backward, sm_100a.
//
// Replaces the ~12 torch kernels behind
//   user_emb, pos, neg = rec_user_emb[user_idx], rec_item_emb[pos_idx], rec_item_emb[neg_idx]   model/graph/LightGCN.py:52
//   bpr_loss(...)      = mean(-log(10e-6 + sigmoid(<u,p> - <u,n>)))                              util/loss_torch.py:5-9
//   l2_reg_loss(reg, u, p, n) / batch_size = reg * (||U_B||_F + ||P_B||_F + ||N_B||_F) / bs      util/loss_torch.py:17-21
// (Frobenius NORMS, not squared; the divisor is the configured batch size even for a short batch.)
//
// One row group of D/4 lanes handles one (user, pos, neg) triple: three coalesced 128-bit row
// gathers, two dot products by shuffle reduction.  Sums over the batch are accumulated in double
// per block and reduced in block order by a second kernel: no float atomics in the forward pass.
// The backward pass scatters row gradients with red.global.add.f32 (rows repeat inside a batch).
#include "hgr_internal.cuh"

namespace hgr {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 1184;

template <int LPR>
__global__ void __launch_bounds__(kLossThreads) bpr_l2_fwd_kernel(const float4 *__restrict__ user_tab,
                                                                  const float4 *__restrict__ item_tab,
                                                                  const int64_t *__restrict__ u, const int64_t *__restrict__ p,
                                                                  const int64_t *__restrict__ n, int64_t batch, int64_t n_users,
                                                                  int64_t n_items, float *__restrict__ saved_x,
                                                                  double *__restrict__ partials, int32_t *__restrict__ bad,
                                                                  int64_t own_lo, int64_t own_hi) {
    // [own_lo, own_hi): only triples whose user row lies in the window are evaluated (owner-sharded training: each rank sums
    // the triples of its own users); saved_x == NULL there.  The default window is everything.
    constexpr int GPB = kLossThreads / LPR;
    const int gl = threadIdx.x % LPR, g = threadIdx.x / LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((threadIdx.x % 32) / LPR * LPR));
    double loss = 0.0, su = 0.0, sp = 0.0, sn = 0.0;
    for (int64_t b = (int64_t)blockIdx.x * GPB + g; b < batch; b += (int64_t)gridDim.x * GPB) {
        const int64_t iu = u[b];
        if (iu >= 0 && iu < n_users && (iu < own_lo || iu >= own_hi)) continue;  // another rank's user
        const int64_t ip = p[b], in = n[b];
        if (iu < 0 || iu >= n_users || ip < 0 || ip >= n_items || in < 0 || in >= n_items) {
            if (gl == 0) atomicAdd(bad, 1);
            if (gl == 0 && saved_x) saved_x[b] = 0.f;
            continue;
        }
        const float4 a = __ldg(user_tab + iu * LPR + gl);
        const float4 q = __ldg(item_tab + ip * LPR + gl);
        const float4 r = __ldg(item_tab + in * LPR + gl);
        const float pos = group_sum<LPR>((a.x * q.x + a.y * q.y) + (a.z * q.z + a.w * q.w), gmask);
        const float neg = group_sum<LPR>((a.x * r.x + a.y * r.y) + (a.z * r.z + a.w * r.w), gmask);
        const float nu = group_sum<LPR>((a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w), gmask);
        const float np_ = group_sum<LPR>((q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w), gmask);
        const float nn = group_sum<LPR>((r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w), gmask);
        if (gl == 0) {
            const float x = pos - neg;
            if (saved_x) saved_x[b] = x;
            const float s = 1.0f / (1.0f + expf(-x));
            loss += (double)(-logf(10e-6f + s));
            su += (double)nu;
            sp += (double)np_;
            sn += (double)nn;
        }
    }
    __shared__ double sh[4][kLossThreads / 1];
    sh[0][threadIdx.x] = loss;
    sh[1][threadIdx.x] = su;
    sh[2][threadIdx.x] = sp;
    sh[3][threadIdx.x] = sn;
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int k = 0; k < GPB; ++k) t += sh[threadIdx.x][k * LPR];  // lane 0 of every group, in group order
        partials[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

__device__ __forceinline__ void bpr_l2_final(const double *t, int64_t batch, float reg, float batch_size_div, float *out, float *norms) {
    const float nu = (float)sqrt(t[1]), np_ = (float)sqrt(t[2]), nn = (float)sqrt(t[3]);
    out[0] = batch > 0 ? (float)(t[0] / (double)batch) : 0.f;
    out[1] = (nu + np_ + nn) * reg / batch_size_div;
    norms[0] = nu;
    norms[1] = np_;
    norms[2] = nn;
}

__global__ void bpr_l2_final_kernel(const double *__restrict__ sums, int64_t batch, float reg, float batch_size_div, float *__restrict__ out,
                                    float *__restrict__ norms) {
    if (threadIdx.x == 0 && blockIdx.x == 0) bpr_l2_final(sums, batch, reg, batch_size_div, out, norms);
}

// out[0] = rec loss, out[1] = reg loss; norms[0..2] = ||U_B||, ||P_B||, ||N_B|| (saved for backward).  With `sums` the four
// totals are written there instead (owner-sharded form: the caller adds them over the ranks first).
__global__ void __launch_bounds__(256) bpr_l2_finish_kernel(const double *__restrict__ partials, int n_blocks, int64_t batch, float reg,
                                                            float batch_size_div, float *__restrict__ out,
                                                            float *__restrict__ norms, double *__restrict__ sums) {
    // thread t adds blocks t, t + 256, ... in order, then a fixed tree: deterministic
    __shared__ double sh[4][256];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < n_blocks; b += 256)
        for (int k = 0; k < 4; ++k) acc[k] += partials[(size_t)b * 4 + k];
    for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w)
            for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const double t[4] = {sh[0][0], sh[1][0], sh[2][0], sh[3][0]};
    if (sums) {
        for (int k = 0; k < 4; ++k) sums[k] = t[k];
        return;
    }
    bpr_l2_final(t, batch, reg, batch_size_div, out, norms);
}

__device__ __forceinline__ void red_add_f4(float *addr, float4 v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v.x) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr + 1), "f"(v.y) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr + 2), "f"(v.z) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr + 3), "f"(v.w) : "memory");
}

template <int LPR>
__global__ void __launch_bounds__(kLossThreads) bpr_l2_bwd_kernel(const float4 *__restrict__ user_tab,
                                                                  const float4 *__restrict__ item_tab,
                                                                  const int64_t *__restrict__ u, const int64_t *__restrict__ p,
                                                                  const int64_t *__restrict__ n, int64_t batch, int64_t n_users,
                                                                  int64_t n_items, const float *__restrict__ saved_x,
                                                                  const float *__restrict__ norms, const float *__restrict__ grad_out,
                                                                  float reg, float batch_size_div, float *__restrict__ d_user,
                                                                  float *__restrict__ d_item, int64_t row_lo, int64_t row_hi) {
    // saved_x == NULL: x is recomputed from the rows (owner-sharded form: the forward of this rank only saw its own users).
    // [row_lo, row_hi): only gradient rows inside the window are produced, at d[row - row_lo] (sharded training: a rank
    // owns a window of the gathered table); the default window is the whole table.
    constexpr int GPB = kLossThreads / LPR;
    const int gl = threadIdx.x % LPR, g = threadIdx.x / LPR;
    const float g_rec = grad_out[0], g_reg = grad_out[1];
    const float c = g_reg * reg / batch_size_div;
    const float cu = norms[0] > 0.f ? c / norms[0] : 0.f;
    const float cp = norms[1] > 0.f ? c / norms[1] : 0.f;
    const float cn = norms[2] > 0.f ? c / norms[2] : 0.f;
    const float inv_b = g_rec / (float)batch;
    for (int64_t b = (int64_t)blockIdx.x * GPB + g; b < batch; b += (int64_t)gridDim.x * GPB) {
        const int64_t iu = u[b], ip = p[b], in = n[b];
        if (iu < 0 || iu >= n_users || ip < 0 || ip >= n_items || in < 0 || in >= n_items) continue;
        const bool wu = iu >= row_lo && iu < row_hi, wp = ip >= row_lo && ip < row_hi, wn = in >= row_lo && in < row_hi;
        if (!(wu || wp || wn)) continue;
        const float4 a = __ldg(user_tab + iu * LPR + gl);
        const float4 q = __ldg(item_tab + ip * LPR + gl);
        const float4 r = __ldg(item_tab + in * LPR + gl);
        float x;
        if (saved_x) {
            x = saved_x[b];
        } else {
            const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((threadIdx.x % 32) / LPR * LPR));
            const float pos = group_sum<LPR>((a.x * q.x + a.y * q.y) + (a.z * q.z + a.w * q.w), gmask);
            const float neg = group_sum<LPR>((a.x * r.x + a.y * r.y) + (a.z * r.z + a.w * r.w), gmask);
            x = pos - neg;
        }
        const float s = 1.0f / (1.0f + expf(-x));
        const float gx = -(s * (1.0f - s)) / (10e-6f + s) * inv_b;  // d loss / d x
        float4 du, dp, dn;
        du.x = gx * (q.x - r.x) + cu * a.x; du.y = gx * (q.y - r.y) + cu * a.y;
        du.z = gx * (q.z - r.z) + cu * a.z; du.w = gx * (q.w - r.w) + cu * a.w;
        dp.x = gx * a.x + cp * q.x; dp.y = gx * a.y + cp * q.y; dp.z = gx * a.z + cp * q.z; dp.w = gx * a.w + cp * q.w;
        dn.x = -gx * a.x + cn * r.x; dn.y = -gx * a.y + cn * r.y; dn.z = -gx * a.z + cn * r.z; dn.w = -gx * a.w + cn * r.w;
        if (wu) red_add_f4(d_user + ((iu - row_lo) * LPR + gl) * 4, du);
        if (wp) red_add_f4(d_item + ((ip - row_lo) * LPR + gl) * 4, dp);
        if (wn) red_add_f4(d_item + ((in - row_lo) * LPR + gl) * 4, dn);
    }
}

static int loss_blocks(int64_t batch, int gpb) {
    int64_t b = ceil_div(batch, gpb);
    if (b < 1) b = 1;
    return (int)(b < kLossMaxBlocks ? b : kLossMaxBlocks);
}

}  // namespace hgr

extern "C" {

size_t hgr_bpr_l2_workspace_bytes(int64_t batch) {
    // saved x per triple + 3 norms (floats), then the per-block double partials, 16-byte aligned
    size_t s = ((size_t)(batch + 4) * sizeof(float) + 15) & ~(size_t)15;
    return s + (size_t)hgr::kLossMaxBlocks * 4 * sizeof(double) + 16;
}

static int bpr_fwd_impl(const float *user_tab, const float *item_tab, int64_t n_users, int64_t n_items, int32_t D,
                        const int64_t *u, const int64_t *p, const int64_t *n, int64_t batch, float reg, float batch_size_div,
                        float *out, void *saved, size_t saved_bytes, int32_t *bad_index_count, int64_t own_lo, int64_t own_hi,
                        double *sums, hgr_stream_t stream) {
    using namespace hgr;
    HGR_REQUIRE(sums || out, "NULL argument");
    HGR_REQUIRE(D == 32 || D == 64 || D == 128, "D = %d unsupported (32, 64 or 128)", D);
    HGR_REQUIRE(batch >= 0 && n_users >= 0 && n_items >= 0, "negative size");
    HGR_REQUIRE(user_tab && item_tab && saved && bad_index_count, "NULL argument");
    HGR_REQUIRE(batch == 0 || (u && p && n), "NULL index array");
    HGR_REQUIRE(aligned16(user_tab) && aligned16(item_tab) && aligned16(saved), "tables and workspace must be 16-byte aligned");
    HGR_REQUIRE(batch_size_div > 0.f, "batch_size_div must be positive");
    if (saved_bytes < hgr_bpr_l2_workspace_bytes(batch))
        return set_error(HGR_ERR_WORKSPACE, "bpr workspace: need %zu bytes, got %zu", hgr_bpr_l2_workspace_bytes(batch), saved_bytes);
    float *saved_x = reinterpret_cast<float *>(saved);
    float *norms = saved_x + batch;
    double *partials = reinterpret_cast<double *>(reinterpret_cast<char *>(saved) + (((size_t)(batch + 4) * 4 + 15) & ~(size_t)15));
    if (sums) saved_x = nullptr;  // owner-sharded: the backward recomputes x
    cudaStream_t st = (cudaStream_t)stream;
    const float4 *ut = reinterpret_cast<const float4 *>(user_tab), *it = reinterpret_cast<const float4 *>(item_tab);
    int blocks;
    switch (D) {
        case 32: blocks = loss_blocks(batch, kLossThreads / 8);
            bpr_l2_fwd_kernel<8><<<blocks, kLossThreads, 0, st>>>(ut, it, u, p, n, batch, n_users, n_items, saved_x, partials, bad_index_count, own_lo, own_hi); break;
        case 64: blocks = loss_blocks(batch, kLossThreads / 16);
            bpr_l2_fwd_kernel<16><<<blocks, kLossThreads, 0, st>>>(ut, it, u, p, n, batch, n_users, n_items, saved_x, partials, bad_index_count, own_lo, own_hi); break;
        default: blocks = loss_blocks(batch, kLossThreads / 32);
            bpr_l2_fwd_kernel<32><<<blocks, kLossThreads, 0, st>>>(ut, it, u, p, n, batch, n_users, n_items, saved_x, partials, bad_index_count, own_lo, own_hi); break;
    }
    HGR_LAUNCH_OK("bpr_l2_fwd_kernel");
    bpr_l2_finish_kernel<<<1, 256, 0, st>>>(partials, blocks, batch, reg, batch_size_div, out, norms, sums);
    HGR_LAUNCH_OK("bpr_l2_finish_kernel");
    return HGR_OK;
}

int hgr_bpr_l2_fwd_f32(const float *user_tab, const float *item_tab, int64_t n_users, int64_t n_items, int32_t D,
                       const int64_t *u, const int64_t *p, const int64_t *n, int64_t batch, float reg, float batch_size_div,
                       float *out, void *saved, size_t saved_bytes, int32_t *bad_index_count, hgr_stream_t stream) {
    HGR_REQUIRE(out, "NULL argument");
    return bpr_fwd_impl(user_tab, item_tab, n_users, n_items, D, u, p, n, batch, reg, batch_size_div, out, saved, saved_bytes,
                        bad_index_count, 0, n_users, nullptr, stream);
}

int hgr_bpr_l2_fwd_owned_f32(const float *table, int64_t n_rows, int32_t D, const int64_t *u, const int64_t *p, const int64_t *n,
                             int64_t batch, int64_t own_lo, int64_t own_hi, double *sums, void *saved, size_t saved_bytes,
                             int32_t *bad_index_count, hgr_stream_t stream) {
    HGR_REQUIRE(sums, "NULL argument");
    HGR_REQUIRE(own_lo >= 0 && own_lo <= own_hi && own_hi <= n_rows, "window [%lld, %lld) outside the table", (long long)own_lo,
                (long long)own_hi);
    return bpr_fwd_impl(table, table, n_rows, n_rows, D, u, p, n, batch, 0.f, 1.f, nullptr, saved, saved_bytes, bad_index_count, own_lo,
                        own_hi, sums, stream);
}

int hgr_bpr_l2_finish_f32(const double *sums, int64_t batch, float reg, float batch_size_div, float *out, float *norms,
                          hgr_stream_t stream) {
    using namespace hgr;
    HGR_REQUIRE(sums && out && norms, "NULL argument");
    HGR_REQUIRE(batch >= 0 && batch_size_div > 0.f, "batch must be >= 0 and batch_size_div positive");
    bpr_l2_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, batch, reg, batch_size_div, out, norms);
    HGR_LAUNCH_OK("bpr_l2_final_kernel");
    return HGR_OK;
}

static int bpr_bwd_impl(const float *user_tab, const float *item_tab, int64_t n_users, int64_t n_items, int32_t D, const int64_t *u,
                        const int64_t *p, const int64_t *n, int64_t batch, float reg, float batch_size_div, const void *saved,
                        const float *grad_out, float *d_user_tab, float *d_item_tab, int64_t row_lo, int64_t row_hi,
                        hgr_stream_t stream, const float *global_norms = nullptr) {
    using namespace hgr;
    HGR_REQUIRE(D == 32 || D == 64 || D == 128, "D = %d unsupported (32, 64 or 128)", D);
    HGR_REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return HGR_OK;
    HGR_REQUIRE(user_tab && item_tab && u && p && n && (saved || global_norms) && grad_out && d_user_tab && d_item_tab, "NULL argument");
    HGR_REQUIRE(aligned16(user_tab) && aligned16(item_tab) && aligned16(d_user_tab) && aligned16(d_item_tab), "tables must be 16-byte aligned");
    const float *saved_x = global_norms ? nullptr : reinterpret_cast<const float *>(saved);
    const float *norms = global_norms ? global_norms : saved_x + batch;
    cudaStream_t st = (cudaStream_t)stream;
    const float4 *ut = reinterpret_cast<const float4 *>(user_tab), *it = reinterpret_cast<const float4 *>(item_tab);
    switch (D) {
        case 32: bpr_l2_bwd_kernel<8><<<loss_blocks(batch, kLossThreads / 8), kLossThreads, 0, st>>>(ut, it, u, p, n, batch, n_users, n_items, saved_x, norms, grad_out, reg, batch_size_div, d_user_tab, d_item_tab, row_lo, row_hi); break;
        case 64: bpr_l2_bwd_kernel<16><<<loss_blocks(batch, kLossThreads / 16), kLossThreads, 0, st>>>(ut, it, u, p, n, batch, n_users, n_items, saved_x, norms, grad_out, reg, batch_size_div, d_user_tab, d_item_tab, row_lo, row_hi); break;
        default: bpr_l2_bwd_kernel<32><<<loss_blocks(batch, kLossThreads / 32), kLossThreads, 0, st>>>(ut, it, u, p, n, batch, n_users, n_items, saved_x, norms, grad_out, reg, batch_size_div, d_user_tab, d_item_tab, row_lo, row_hi); break;
    }
    HGR_LAUNCH_OK("bpr_l2_bwd_kernel");
    return HGR_OK;
}

int hgr_bpr_l2_bwd_f32(const float *user_tab, const float *item_tab, int64_t n_users, int64_t n_items, int32_t D,
                       const int64_t *u, const int64_t *p, const int64_t *n, int64_t batch, float reg, float batch_size_div,
                       const void *saved, const float *grad_out, float *d_user_tab, float *d_item_tab, hgr_stream_t stream) {
    const int64_t hi = n_users > n_items ? n_users : n_items;
    return bpr_bwd_impl(user_tab, item_tab, n_users, n_items, D, u, p, n, batch, reg, batch_size_div, saved, grad_out, d_user_tab,
                        d_item_tab, 0, hi, stream);
}

int hgr_bpr_l2_bwd_window_f32(const float *table, int64_t n_rows, int32_t D, const int64_t *u, const int64_t *p, const int64_t *n,
                              int64_t batch, float reg, float batch_size_div, const void *saved, const float *grad_out,
                              int64_t row_lo, int64_t row_hi, float *d_rows, hgr_stream_t stream) {
    HGR_REQUIRE(row_lo >= 0 && row_lo <= row_hi && row_hi <= n_rows, "window [%lld, %lld) outside the table", (long long)row_lo,
                (long long)row_hi);
    return bpr_bwd_impl(table, table, n_rows, n_rows, D, u, p, n, batch, reg, batch_size_div, saved, grad_out, d_rows, d_rows, row_lo,
                        row_hi, stream);
}

int hgr_bpr_l2_bwd_owned_f32(const float *table, int64_t n_rows, int32_t D, const int64_t *u, const int64_t *p, const int64_t *n,
                             int64_t batch, float reg, float batch_size_div, const float *norms, const float *grad_out,
                             int64_t row_lo, int64_t row_hi, float *d_rows, hgr_stream_t stream) {
    HGR_REQUIRE(norms, "NULL argument");
    HGR_REQUIRE(row_lo >= 0 && row_lo <= row_hi && row_hi <= n_rows, "window [%lld, %lld) outside the table", (long long)row_lo,
                (long long)row_hi);
    return bpr_bwd_impl(table, table, n_rows, n_rows, D, u, p, n, batch, reg, batch_size_div, nullptr, grad_out, d_rows, d_rows, row_lo,
                        row_hi, stream, norms);
}

}  // extern "C"
