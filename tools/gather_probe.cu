// Microbenchmark, not product code: what can this B200 deliver for the access pattern of the propagation kernel?
//
//   gp_gather  256-byte (D = 64 fp32) row gathers X[idx[p]] for a stream of column ids, NO CSR walk, NO per-row
//              structure, NO multiply: every 16-lane group takes an equal slice of the id stream, keeps a ring of row
//              copies in flight (cp.async, like spmm_rows_async_kernel) or a batch of register loads, and adds the rows up.
//              Fed with the indices array of the benchmark graph it sees exactly the column distribution (and L2 hit
//              rate) of the product kernel, so its time is the floor for "gather these rows at all".
//   gp_stream  coalesced 128-bit reads of a buffer, `repeat` passes: an L2-resident buffer gives the L2 -> SM fabric
//              peak, a large one the DRAM read peak.
//
// Built by tools/gather_probe.py with nvcc -gencode arch=compute_100a,code=sm_100a into tools/libgather_probe.so.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kThreads = 256;
constexpr int LPR = 16;  // lanes per 256-byte row

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ int ld_stream_i32(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ring of STAGES sub-batches of HB rows per group; STAGES - 1 sub-batches in flight while one is consumed
template <int HB, int STAGES, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) gather_ring_kernel(const int *__restrict__ idx, int64_t nnz, const float4 *__restrict__ X4,
                                                                 float4 *__restrict__ out, int per_group) {
    static_assert(LPR % HB == 0, "a sub-batch must not straddle two id batches");
    extern __shared__ float4 ring_all[];  // [STAGES * HB][kThreads]
    float4 *ring = ring_all + threadIdx.x;
    const int gl = threadIdx.x % LPR;
    const unsigned gmask = 0xffffu << ((threadIdx.x % 32) / LPR * LPR);
    const int64_t group = (int64_t)blockIdx.x * (kThreads / LPR) + threadIdx.x / LPR;
    const int64_t s = group * per_group;
    int64_t e = s + per_group;
    if (e > nnz) e = nnz;
    if (s >= e) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // per_group is a multiple of LPR: batch b = ids [s + b * LPR, +LPR), one per lane
    const int n_sub = (int)((e - s + HB - 1) / HB);
    int c_iss = (s + gl < e) ? ld_stream_i32(idx + s + gl) : 0;              // ids of the batch being issued, one per lane
    int c_pre = (s + LPR + gl < e) ? ld_stream_i32(idx + s + LPR + gl) : 0;  // the batch after it (prefetched)
#pragma unroll 1
    for (int sub = 0; sub < n_sub + STAGES - 1; ++sub) {
        if (sub < n_sub) {
            const int first = (sub * HB) % LPR;  // position of the sub-batch inside its batch of LPR ids
#pragma unroll
            for (int k = 0; k < HB; ++k) {
                const int cc = __shfl_sync(gmask, c_iss, first + k, LPR);
                if (s + (int64_t)sub * HB + k < e) cp_async16(ring + ((sub % STAGES) * HB + k) * kThreads, X4 + (int64_t)cc * LPR + gl);
            }
            if (first + HB == LPR) {
                c_iss = c_pre;
                const int64_t p = s + ((int64_t)sub * HB / LPR + 2) * LPR + gl;
                c_pre = (p < e) ? ld_stream_i32(idx + p) : 0;
            }
        }
        cp_async_commit();
        const int cons = sub - (STAGES - 1);
        if (cons >= 0) {
            cp_async_wait<STAGES - 1>();
#pragma unroll
            for (int k = 0; k < HB; ++k) {
                if (s + (int64_t)cons * HB + k < e) {
                    const float4 x = ring[((cons % STAGES) * HB + k) * kThreads];
                    acc.x += x.x;
                    acc.y += x.y;
                    acc.z += x.z;
                    acc.w += x.w;
                }
            }
        }
    }
    out[group * LPR + gl] = acc;
}

// register gathers: UNR independent loads per group in flight
template <int UNR, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) gather_reg_kernel(const int *__restrict__ idx, int64_t nnz, const float4 *__restrict__ X4,
                                                                float4 *__restrict__ out, int per_group) {
    const int gl = threadIdx.x % LPR;
    const unsigned gmask = 0xffffu << ((threadIdx.x % 32) / LPR * LPR);
    const int64_t group = (int64_t)blockIdx.x * (kThreads / LPR) + threadIdx.x / LPR;
    const int64_t s = group * per_group;
    int64_t e = s + per_group;
    if (e > nnz) e = nnz;
    if (s >= e) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = (s + gl < e) ? ld_stream_i32(idx + s + gl) : 0;
    for (int64_t base = s; base < e; base += LPR) {
        const int cn = (base + LPR + gl < e) ? ld_stream_i32(idx + base + LPR + gl) : 0;
#pragma unroll
        for (int k0 = 0; k0 < LPR; k0 += UNR) {
            float4 x[UNR];
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                const int cc = __shfl_sync(gmask, c, k0 + k, LPR);
                x[k] = (base + k0 + k < e) ? __ldg(X4 + (int64_t)cc * LPR + gl) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                acc.x += x[k].x;
                acc.y += x[k].y;
                acc.z += x[k].z;
                acc.w += x[k].w;
            }
        }
        c = cn;
    }
    out[group * LPR + gl] = acc;
}

// TMA tile::gather4 (sm_100): ONE instruction from one lane fetches the 4 table rows whose indices it names (1 KB) into
// shared memory and signals an mbarrier; no per-lane LDGSTS.  Ring of STAGES x 4 rows per 16-lane group, one mbarrier per stage.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (!done && spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *tm, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
                 : "memory");
}

template <int STAGES, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) gather_tma_kernel(const int *__restrict__ idx, int64_t nnz, const __grid_constant__ CUtensorMap tm,
                                                                float4 *__restrict__ out, int per_group) {
    extern __shared__ __align__(1024) uint8_t tma_smem[];  // [groups][STAGES][4 rows][256 B], then the mbarriers
    constexpr int GPB = kThreads / LPR;
    const int g = threadIdx.x / LPR, gl = threadIdx.x % LPR;
    const unsigned gmask = 0xffffu << ((threadIdx.x % 32) / LPR * LPR);
    uint8_t *ring = tma_smem + (size_t)g * STAGES * 1024;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tma_smem + (size_t)GPB * STAGES * 1024) + g * STAGES;
    if (gl == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_addr(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int64_t group = (int64_t)blockIdx.x * GPB + g;
    const int64_t s0 = group * per_group;
    int64_t e = s0 + per_group;
    if (e > nnz) e = nnz;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s0 < e) {
        const int n_sub = (int)((e - s0) / 4);  // per_group is a multiple of 16: whole quads (the tail of the stream is dropped)
        int c_iss = (s0 + gl < e) ? ld_stream_i32(idx + s0 + gl) : 0;
        int c_pre = (s0 + LPR + gl < e) ? ld_stream_i32(idx + s0 + LPR + gl) : 0;
#pragma unroll 1
        for (int sub = 0; sub < n_sub + STAGES - 1; ++sub) {
            if (sub < n_sub) {
                const int first = (sub * 4) % LPR;
                const int r0 = __shfl_sync(gmask, c_iss, first, LPR), r1 = __shfl_sync(gmask, c_iss, first + 1, LPR);
                const int r2 = __shfl_sync(gmask, c_iss, first + 2, LPR), r3 = __shfl_sync(gmask, c_iss, first + 3, LPR);
                if (gl == 0) {
                    const uint32_t bar = smem_addr(bars + sub % STAGES);
                    mbar_expect(bar, 1024);
                    tma_gather4(smem_addr(ring + (sub % STAGES) * 1024), &tm, 0, r0, r1, r2, r3, bar);
                }
                if (first + 4 == LPR) {
                    c_iss = c_pre;
                    const int64_t p = s0 + ((int64_t)sub * 4 / LPR + 2) * LPR + gl;
                    c_pre = (p < e) ? ld_stream_i32(idx + p) : 0;
                }
            }
            const int cons = sub - (STAGES - 1);
            if (cons >= 0) {
                mbar_wait(smem_addr(bars + cons % STAGES), (cons / STAGES) & 1);
                const float4 *rows = reinterpret_cast<const float4 *>(ring + (cons % STAGES) * 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 x = rows[k * LPR + gl];
                    acc.x += x.x;
                    acc.y += x.y;
                    acc.z += x.z;
                    acc.w += x.w;
                }
                __syncwarp(gmask);  // every lane of the group has read the stage before lane 0 re-arms it
            }
        }
    }
    out[group * LPR + gl] = acc;
}

__global__ void __launch_bounds__(kThreads) stream_kernel(const float4 *__restrict__ buf, int64_t n4, int repeat, float4 *__restrict__ out) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int r = 0; r < repeat; ++r) {
        int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 a, b, c, d;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(buf + i));
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(buf + i + stride));
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(buf + i + 2 * stride));
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(buf + i + 3 * stride));
            acc.x += (a.x + b.x) + (c.x + d.x);
            acc.y += (a.y + b.y) + (c.y + d.y);
            acc.z += (a.z + b.z) + (c.z + d.z);
            acc.w += (a.w + b.w) + (c.w + d.w);
        }
        for (; i < n4; i += stride) {
            float4 a;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(buf + i));
            acc.x += a.x;
            acc.y += a.y;
            acc.z += a.z;
            acc.w += a.w;
        }
    }
    if (acc.x == 123.456f) out[0] = acc;  // keep the loads alive
}

template <int HB, int STAGES, int MINB>
int launch_ring(const int *idx, int64_t nnz, const float *X, float *out, int per_group, cudaStream_t st) {
    const int64_t groups = (nnz + per_group - 1) / per_group;
    const int64_t grid = (groups + kThreads / LPR - 1) / (kThreads / LPR);
    const size_t smem = (size_t)STAGES * HB * kThreads * sizeof(float4);
    cudaError_t e = cudaFuncSetAttribute(gather_ring_kernel<HB, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -2;
    gather_ring_kernel<HB, STAGES, MINB><<<(unsigned)grid, kThreads, smem, st>>>(idx, nnz, (const float4 *)X, (float4 *)out, per_group);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

template <int UNR, int MINB>
int launch_reg(const int *idx, int64_t nnz, const float *X, float *out, int per_group, cudaStream_t st) {
    const int64_t groups = (nnz + per_group - 1) / per_group;
    const int64_t grid = (groups + kThreads / LPR - 1) / (kThreads / LPR);
    gather_reg_kernel<UNR, MINB><<<(unsigned)grid, kThreads, 0, st>>>(idx, nnz, (const float4 *)X, (float4 *)out, per_group);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace

extern "C" {

// out: [ceil(nnz / per_group), 64] floats.  per_group: multiple of 16.
int gp_gather(const int *idx, int64_t nnz, const float *X, float *out, int variant, int per_group, cudaStream_t st) {
    if (per_group <= 0 || per_group % LPR) return -1;
    switch (variant) {
        case 0: return launch_ring<4, 2, 5>(idx, nnz, X, out, per_group, st);   // the product kernel's ring: 4 in flight, 5 blocks/SM
        case 1: return launch_ring<4, 3, 4>(idx, nnz, X, out, per_group, st);   // 8 in flight, 4 blocks
        case 2: return launch_ring<4, 4, 3>(idx, nnz, X, out, per_group, st);   // 12 in flight, 3 blocks
        case 3: return launch_ring<8, 2, 3>(idx, nnz, X, out, per_group, st);   // 8 in flight, 3 blocks
        case 4: return launch_ring<8, 3, 2>(idx, nnz, X, out, per_group, st);   // 16 in flight, 2 blocks
        case 5: return launch_ring<2, 4, 6>(idx, nnz, X, out, per_group, st);   // 6 in flight, 6 blocks
        case 6: return launch_ring<2, 2, 8>(idx, nnz, X, out, per_group, st);   // 2 in flight, 8 blocks
        case 7: return launch_ring<4, 2, 7>(idx, nnz, X, out, per_group, st);
        case 10: return launch_reg<4, 6>(idx, nnz, X, out, per_group, st);
        case 11: return launch_reg<8, 4>(idx, nnz, X, out, per_group, st);
        case 12: return launch_reg<16, 2>(idx, nnz, X, out, per_group, st);
        default: return -1;
    }
}

// TMA gather4 variant: table [n_rows, 64] fp32.  Returns -3 when the driver refuses the tensor map.
int gp_gather_tma(const int *idx, int64_t nnz, const float *X, int64_t n_rows, float *out, int variant, int per_group, int box_rows, cudaStream_t st) {
    if (per_group <= 0 || per_group % LPR) return -1;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -3;
    CUtensorMap tm;
    const cuuint64_t dims[2] = {64, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {256};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -3;
    const int64_t groups = (nnz + per_group - 1) / per_group;
    const int64_t grid = (groups + kThreads / LPR - 1) / (kThreads / LPR);
#define GP_TMA(STG, MB)                                                                                                        \
    do {                                                                                                                       \
        const size_t smem = (size_t)(kThreads / LPR) * STG * 1024 + (size_t)(kThreads / LPR) * STG * 8;                         \
        if (cudaFuncSetAttribute(gather_tma_kernel<STG, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2; \
        gather_tma_kernel<STG, MB><<<(unsigned)grid, kThreads, smem, st>>>(idx, nnz, tm, (float4 *)out, per_group);             \
    } while (0)
    switch (variant) {
        case 0: GP_TMA(2, 5); break;
        case 1: GP_TMA(3, 4); break;
        case 2: GP_TMA(4, 3); break;
        case 3: GP_TMA(2, 6); break;
        default: return -1;
    }
#undef GP_TMA
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int gp_stream(const float *buf, int64_t n_floats, int repeat, int blocks, float *out, cudaStream_t st) {
    stream_kernel<<<blocks, kThreads, 0, st>>>((const float4 *)buf, n_floats / 4, repeat, (float4 *)out);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
