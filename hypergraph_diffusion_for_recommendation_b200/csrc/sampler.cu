// BPR negative sampling on the device, sm_100a.
//
// Replaces the body of next_batch_pairwise (util/sampler.py:237-264): for every positive (user, item) of the batch
// draw items uniformly from the whole catalogue until one is not in the user's training set (rejection sampling;
// `choice(item_list)` ... `while neg_item in data.training_set_u[user]`).  The reference does this in python, one
// `random.choice` at a time, and rebuilds `item_list` for every batch (0.1-0.7 s per batch, impossible
// at 1 B interactions).  Here one thread per (positive, negative slot): Philox4x32-10 keyed by the seed with the
// counter (sample index, attempt), membership by binary search in the user's sorted training row.
// The random stream is NOT python's Mersenne Twister: parity runs replay the reference's triples instead (the loss
// kernels take any index tensors); this kernel is for throughput runs and is checked for its distribution.
#include "hgr_internal.cuh"

namespace hgr {

// unbiased integer in [0, n): 64-bit multiply-shift of a 32-bit draw with rejection of the short tail (Lemire)
__device__ __forceinline__ bool bounded(uint32_t x, uint32_t n, uint32_t &r) {
    const uint64_t m = (uint64_t)x * n;
    const uint32_t l = (uint32_t)m;
    if (l < n) {
        const uint32_t t = (0u - n) % n;
        if (l < t) return false;
    }
    r = (uint32_t)(m >> 32);
    return true;
}

constexpr int kSampleMaxRounds = 64;  // 4 draws per round; after 256 rejected draws the last draw is kept

__global__ void __launch_bounds__(256) bpr_sample_kernel(const int32_t *__restrict__ edge_u, const int32_t *__restrict__ edge_i,
                                                         const int64_t *__restrict__ perm, int64_t perm_offset, int64_t batch,
                                                         int32_t n_negs, const int64_t *__restrict__ train_indptr,
                                                         const int32_t *__restrict__ train_indices, int32_t n_items, uint64_t seed,
                                                         uint64_t stream_offset, int64_t *__restrict__ out_u, int64_t *__restrict__ out_p,
                                                         int64_t *__restrict__ out_n, int32_t *__restrict__ gave_up) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * n_negs) return;
    const int64_t b = t / n_negs;
    const int64_t e = perm ? perm[perm_offset + b] : perm_offset + b;
    const int32_t u = edge_u[e];
    if (t % n_negs == 0) {
        out_u[b] = u;
        out_p[b] = edge_i[e];
    }
    const int64_t lo0 = train_indptr[u], hi0 = train_indptr[u + 1];
    uint32_t neg = 0;
    bool found = false;
    for (int round = 0; round < kSampleMaxRounds && !found; ++round) {
        uint32_t r[4];
        philox4x32_10(seed, stream_offset + (uint64_t)t, (uint32_t)round, r);
#pragma unroll
        for (int j = 0; j < 4 && !found; ++j) {
            uint32_t cand;
            if (!bounded(r[j], (uint32_t)n_items, cand)) continue;
            neg = cand;
            int64_t lo = lo0, hi = hi0;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(train_indices + mid) < (int32_t)cand) lo = mid + 1;
                else hi = mid;
            }
            found = !(lo < hi0 && __ldg(train_indices + lo) == (int32_t)cand);
        }
    }
    if (!found && gave_up) atomicAdd(gave_up, 1);
    out_n[t] = neg;  // layout [batch][n_negs], the order the reference appends to j_idx
}

}  // namespace hgr

extern "C" {

int hgr_bpr_sample(const int32_t *edge_u, const int32_t *edge_i, int64_t n_edges, const int64_t *perm, int64_t perm_offset,
                   int64_t batch, int32_t n_negs, const int64_t *train_indptr, const int32_t *train_indices, int32_t n_items,
                   uint64_t seed, uint64_t stream_offset, int64_t *out_u, int64_t *out_p, int64_t *out_n, int32_t *gave_up,
                   hgr_stream_t stream) {
    using namespace hgr;
    HGR_REQUIRE(batch >= 0 && n_negs >= 1 && n_items >= 1, "bad batch / n_negs / n_items");
    HGR_REQUIRE(perm_offset >= 0 && perm_offset + batch <= n_edges, "batch [%lld, %lld) outside the %lld training pairs",
                (long long)perm_offset, (long long)(perm_offset + batch), (long long)n_edges);
    if (batch == 0) return HGR_OK;
    HGR_REQUIRE(edge_u && edge_i && train_indptr && out_u && out_p && out_n, "NULL argument");
    const int64_t n = batch * n_negs;
    bpr_sample_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(edge_u, edge_i, perm, perm_offset, batch, n_negs,
                                                                                  train_indptr, train_indices, n_items, seed,
                                                                                  stream_offset, out_u, out_p, out_n, gave_up);
    HGR_LAUNCH_OK("bpr_sample_kernel");
    return HGR_OK;
}

}  // extern "C"
