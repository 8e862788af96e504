// Bernoulli edge dropout of a sparse matrix, sm_100a.
//
// Replaces SpAdjDropEdge.forward (model/graph/HCCF.py:217-226, the same class in HGNN_HD3.py / HGNN_HD4.py / HCCF_diffusion.py):
//     mask    = ((torch.rand(nnz) + keepRate).floor()).type(torch.bool)        # CPU random numbers, every layer of every batch
//     newVals = vals[mask] / keepRate;  newIdxs = idxs[:, mask]
//     return torch.sparse.FloatTensor(newIdxs, newVals, adj.shape)             # a new COO tensor, uploaded again
// Here the sparsity pattern (and with it the split plan and the schedule of the propagation kernel) stays; a dropped entry
// becomes an explicit zero, which leaves every row sum bit-identical to the compacted matrix (fma(0, x, acc) == acc).
// One thread per stored entry:
//   out[p] = keep(p) ? values[p] / keep_rate : 0            (a true IEEE division, like the reference's)
//   keep(p) = floor(u + keep_rate) != 0 in fp32, u =
//       rand[p]                                   the caller's uniform numbers (replay of the reference's CPU stream), or
//       Philox4x32-10(seed, row * n_cols + col)   24-bit uniform keyed by the entry's COORDINATES
// mirror != 0 produces the values of the TRANSPOSED dropped matrix on the same (structurally symmetric) pattern -- what the
// backward propagation needs -- by keying entry (r, c) with the coordinates (c, r) (Philox), or by reading rand[rand_pos[p]]
// where rand_pos is the transpose permutation (replay mode).  No sort, no compaction, no host round trip.
#include "hgr_internal.cuh"

namespace hgr {

__global__ void __launch_bounds__(256) drop_edges_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                         const float *__restrict__ values, int32_t n_rows, int64_t n_cols, int64_t nnz,
                                                         float keep, uint64_t seed, const uint64_t *__restrict__ seed_dev,
                                                         const float *__restrict__ rand, const int64_t *__restrict__ rand_pos, int mirror,
                                                         float *__restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    float u;
    if (rand) {
        u = rand[rand_pos ? rand_pos[p] : p];
    } else {
        // row of entry p: the last row whose offset is <= p
        int lo = 0, hi = n_rows;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (indptr[mid] <= p) lo = mid;
            else hi = mid;
        }
        const int64_t r = lo, c = indices[p];
        const uint64_t key = mirror ? (uint64_t)c * (uint64_t)n_cols + (uint64_t)r : (uint64_t)r * (uint64_t)n_cols + (uint64_t)c;
        uint32_t x[4];
        // seed_dev: a step counter that lives on the device, so that a CUDA graph which captured this launch draws a fresh
        // mask at every replay
        philox4x32_10(seed_dev ? seed + *seed_dev * 0x9E3779B97F4A7C15ull : seed, key, 0u, x);
        u = (float)(x[0] >> 8) * (1.0f / 16777216.0f);  // 24 random bits: [0, 1) like torch.rand's float32
    }
    const bool kept = floorf(u + keep) != 0.f;
    out[p] = kept ? __fdiv_rn(values[p], keep) : 0.f;
}

}  // namespace hgr

extern "C" int hgr_drop_edges_f32(const int64_t *indptr, const int32_t *indices, const float *values, int32_t n_rows, int64_t n_cols,
                                  int64_t nnz, float keep, uint64_t seed, const uint64_t *seed_dev, const float *rand, const int64_t *rand_pos,
                                  int32_t mirror, float *out, hgr_stream_t stream) {
    using namespace hgr;
    HGR_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "negative dimension");
    HGR_REQUIRE(keep > 0.f && keep <= 1.f, "keep rate %g outside (0, 1]", (double)keep);
    HGR_REQUIRE(rand_pos == nullptr || rand != nullptr, "rand_pos without rand");
    if (nnz == 0) return HGR_OK;
    HGR_REQUIRE(indptr && indices && values && out, "NULL argument");
    drop_edges_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, (cudaStream_t)stream>>>(indptr, indices, values, n_rows, n_cols, nnz, keep, seed,
                                                                                    seed_dev, rand, rand_pos, mirror, out);
    HGR_LAUNCH_OK("drop_edges_kernel");
    return HGR_OK;
}
