/*
 * hgr_oracle.c -- CPU restatement of the reference's hot-path arithmetic.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's shared object; the product path (libhgr.so) never does.
 *
 * The reference (DanbiAubrey/Hypergraph_diffusion_for_recommendation, /root/reference/HD_SELFRec)
 * is pure Python; its arithmetic on this path lives in PyTorch / numpy / numba calls.  Each
 * function below restates the published behaviour of one such call site and is pinned against
 * tests/golden/reference_vectors.npz, which was produced by running the reference itself
 * (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -mfma -ffp-contract=off -fPIC -shared (oracle/build.py).  fmaf() must compile to a
 * single-rounding fused multiply-add; every other expression keeps separate roundings.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* torch.sparse.mm(A, X) on CPU (call sites: model/graph/LightGCN.py:133, HCCF.py:199,
 * HGNN_HD3.py:549-553): for every row the nonzeros are visited in stored order and each output
 * feature is accumulated with one fused multiply-add per nonzero (probed). */
void hgr_oracle_spmm_f32(const int64_t *indptr, const int64_t *indices, const float *data,
                         int64_t n_rows, int64_t d, const float *x, float *y) {
    for (int64_t r = 0; r < n_rows; ++r) {
        float *yr = y + r * d;
        for (int64_t k = 0; k < d; ++k) yr[k] = 0.0f;
        for (int64_t p = indptr[r]; p < indptr[r + 1]; ++p) {
            const float v = data[p];
            const float *xc = x + indices[p] * d;
            for (int64_t k = 0; k < d; ++k) yr[k] = fmaf(v, xc[k], yr[k]);
        }
    }
}

/* Canonical full-ranking score: s[u][i] = sum_k user[u][k] * item[i][k], accumulated with fmaf in
 * ascending k (restates torch.matmul(user_emb[u], item_emb.T), model/graph/LightGCN.py:99-102; the
 * BLAS summation order of the reference is unspecified, so the ascending-k fused order is the
 * definition both the oracle and the CUDA path use for exact top-K comparison). */
void hgr_oracle_scores_f32(const float *user, const float *item, int64_t n_items, int64_t d, float *out) {
    for (int64_t i = 0; i < n_items; ++i) {
        float acc = 0.0f;
        const float *it = item + i * d;
        for (int64_t k = 0; k < d; ++k) acc = fmaf(user[k], it[k], acc);
        out[i] = acc;
    }
}

/* numba's list.sort(key=..., reverse=True) (numba/cpython/listobj.py ol_list_sort ->
 * numba/misc/quicksort.py; pinned numba==0.53.1 in the reference's requirements.txt, 0.65.0 in the
 * build container, same algorithm): an ARGSORT quicksort whose "less than" is `a > b`, median-of-3
 * pivot, partitions of 15 or fewer elements finished by insertion sort, larger side pushed on a
 * stack.  Restated here because the order of exactly tied scores among the first K candidates is
 * whatever this procedure leaves (it is not a stable sort once len > 15). */
#define NB_SMALL 15
static inline int nb_lt(float a, float b) { return a > b; }

static void nb_insertion(const float *a, int64_t *r, int64_t low, int64_t high) {
    for (int64_t i = low + 1; i <= high; ++i) {
        int64_t k = r[i], j = i;
        float v = a[k];
        while (j > low && nb_lt(v, a[r[j - 1]])) {
            r[j] = r[j - 1];
            --j;
        }
        r[j] = k;
    }
}

#define NB_SWAP(x, y) do { int64_t t_ = r[x]; r[x] = r[y]; r[y] = t_; } while (0)
static int64_t nb_partition(const float *a, int64_t *r, int64_t low, int64_t high) {
    int64_t mid = (low + high) >> 1;
    if (nb_lt(a[r[mid]], a[r[low]])) NB_SWAP(low, mid);
    if (nb_lt(a[r[high]], a[r[mid]])) NB_SWAP(high, mid);
    if (nb_lt(a[r[mid]], a[r[low]])) NB_SWAP(low, mid);
    float pivot = a[r[mid]];
    NB_SWAP(high, mid);
    int64_t i = low, j = high - 1;
    for (;;) {
        while (i < high && nb_lt(a[r[i]], pivot)) ++i;
        while (j >= low && nb_lt(pivot, a[r[j]])) --j;
        if (i >= j) break;
        NB_SWAP(i, j);
        ++i;
        --j;
    }
    NB_SWAP(i, high);
    return i;
}

void hgr_oracle_numba_argsort_desc(const float *a, int64_t n, int64_t *r) {
    for (int64_t i = 0; i < n; ++i) r[i] = i;
    if (n < 2) return;
    int64_t st_lo[128], st_hi[128], sp = 1;
    st_lo[0] = 0;
    st_hi[0] = n - 1;
    while (sp > 0) {
        --sp;
        int64_t low = st_lo[sp], high = st_hi[sp];
        while (high - low >= NB_SMALL) {
            int64_t i = nb_partition(a, r, low, high);
            if (high - i > i - low) {
                if (high > i) { st_lo[sp] = i + 1; st_hi[sp] = high; ++sp; }
                high = i - 1;
            } else {
                if (i > low) { st_lo[sp] = low; st_hi[sp] = i - 1; ++sp; }
                low = i + 1;
            }
        }
        nb_insertion(a, r, low, high);
    }
}

/* util/algorithm.py:143-173 find_k_largest, restated INCLUDING the re-visit of the first K
 * candidates.  The initial list is candidates[0:K] in numba's reverse-sorted order;
 * insertion requires strictly greater than the current K-th score and lands after all entries
 * with score >= the new one (the binary search at :154-165). */
void hgr_oracle_find_k_largest(int64_t K, const float *cand, int64_t n, int64_t *ids, float *scores) {
    hgr_oracle_numba_argsort_desc(cand, K, ids);
    for (int64_t j = 0; j < K; ++j) scores[j] = cand[ids[j]];
    for (int64_t iid = 0; iid < n; ++iid) {
        float s = cand[iid];
        if (!(scores[K - 1] < s)) continue;
        int64_t pos = 0;
        while (pos < K && scores[pos] >= s) ++pos;
        for (int64_t j = K - 1; j > pos; --j) {
            scores[j] = scores[j - 1];
            ids[j] = ids[j - 1];
        }
        scores[pos] = s;
        ids[pos] = iid;
    }
}

/* True top-K: score descending, ties by ascending item id, no duplicates ("exact" mode of the
 * drop-in).  Entries equal to mask_value are still candidates, exactly as
 * the reference leaves -10e8 entries in the candidate vector (base/graph_recommender.py:78-83). */
void hgr_oracle_topk_exact(int64_t K, const float *cand, int64_t n, int64_t *ids, float *scores) {
    int64_t filled = 0;
    for (int64_t iid = 0; iid < n; ++iid) {
        float s = cand[iid];
        if (filled == K && !(s > scores[K - 1])) continue;
        int64_t pos = filled < K ? filled : K - 1;
        while (pos > 0 && scores[pos - 1] < s) {
            if (pos < K) {
                scores[pos] = scores[pos - 1];
                ids[pos] = ids[pos - 1];
            }
            --pos;
        }
        scores[pos] = s;
        ids[pos] = iid;
        if (filled < K) ++filled;
    }
}
