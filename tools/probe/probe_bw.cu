// probe_bw.cu -- three B200 measurements the conv kernel's design depends on (run on the GPU box, prints a table):
//   1. TMA L2 -> shared-memory bandwidth: every CTA streaming its own L2-resident region / all CTAs the same region /
//      an HBM-sized region, for 37..444 CTAs -- is the ~9.5 TB/s ceiling seen by the conv kernel per SM or chip-wide,
//      and do loads of the SAME tile by many SMs cost less?
//   2. tcgen05.mma issue rate from shared-memory operands (kind::f16, M = 128, N = 64 / 128 / 256), 1..3 CTAs per SM:
//      is the N = 64 tile limited by shared-memory operand reads?
//   3. SiLU epilogue arithmetic: ex2 + rcp (2 MUFU) vs tanh.approx.f32 (1 MUFU) vs tanh.approx.f16x2 -- elements / clk / SM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ------------------------------------------------------------------------------------------------ 1. TMA bandwidth
constexpr int kStages = 6;
constexpr int kBoxRows = 128;                   // 128 rows x 128 B = 16 KB per box
// second experiment: P producer warps per CTA (each with its own stages and barriers), boxes of `box_rows` rows
__global__ void __launch_bounds__(128) tma_probe_multi(const __grid_constant__ CUtensorMap map, int iters, int box_rows, int producers, int stages,
                                                       long long total_rows, unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[4][8];
    if (threadIdx.x == 0) {
        for (int w = 0; w < 4; ++w) for (int s = 0; s < 8; ++s) mbar_init(&full[w][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < producers) {
        const int box_bytes = box_rows * 128;
        uint8_t* mine = smem + w * stages * box_bytes;
        const long long base = ((long long)(blockIdx.x * 4 + w) * 2048) % (total_rows - 4096);       // own 256 KB region per producer
        const int boxes = 2048 / box_rows;
        const unsigned long long t0 = clock64();
        for (int i = 0; i < iters + stages; ++i) {
            const int s = i % stages;
            if (i >= stages) mbar_wait(&full[w][s], (uint32_t)(((i / stages) - 1) & 1));
            if (i < iters) {
                mbar_expect_tx(&full[w][s], box_bytes);
                const int row = (int)(base + (long long)(i % boxes) * box_rows);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(mine + s * box_bytes)), "l"(&map), "r"(smem_u32(&full[w][s])), "r"(0), "r"(row) : "memory");
            }
        }
        if (w == 0) cycles[blockIdx.x] = clock64() - t0;
    }
}

// third experiment: 3-D boxes {64 elements, rows, chunks} (chunk stride 128 B: [row][chunk][64] in memory -> [chunk][row][64] in shared
// memory): how does the per-CTA fill rate scale with the BOX size (one producer, `stages` boxes in flight)?
__global__ void __launch_bounds__(128) tma_probe_3d(const __grid_constant__ CUtensorMap map, int iters, int box_bytes, int rows, int stages,
                                                    int region_rows, unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int boxes = region_rows / rows;
        const int base = (int)(blockIdx.x * region_rows);
        const unsigned long long t0 = clock64();
        for (int i = 0; i < iters + stages; ++i) {
            const int s = i % stages;
            if (i >= stages) mbar_wait(&full[s], (uint32_t)(((i / stages) - 1) & 1));
            if (i < iters) {
                mbar_expect_tx(&full[s], box_bytes);
                const int row = base + ((i + blockIdx.x) % boxes) * rows;
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(smem + s * box_bytes)), "l"(&map), "r"(smem_u32(&full[s])), "r"(0), "r"(row), "r"(0) : "memory");
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

__global__ void __launch_bounds__(128) tma_probe(const __grid_constant__ CUtensorMap map, int iters, int region_rows, int shared_region,
                                                 long long total_rows, unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[kStages];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // shared_region: 0 = every CTA its own region, 1 = all CTAs the same region, 2 = CTA pairs share a region
        const long long owner = shared_region == 1 ? 0 : (shared_region == 2 ? blockIdx.x / 2 : blockIdx.x);
        const long long base = (owner * (long long)region_rows) % (total_rows - region_rows + 1);
        const int boxes = region_rows / kBoxRows;
        const unsigned long long t0 = clock64();
        for (int i = 0; i < iters + kStages; ++i) {
            const int s = i % kStages;
            if (i >= kStages) mbar_wait(&full[s], (uint32_t)(((i / kStages) - 1) & 1));
            if (i < iters) {
                mbar_expect_tx(&full[s], kBoxRows * 128);
                const int row = (int)(base + (long long)((i + (shared_region ? 0 : blockIdx.x)) % boxes) * kBoxRows);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(smem + s * kBoxRows * 128)), "l"(&map), "r"(smem_u32(&full[s])), "r"(0), "r"(row) : "memory");
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

// ------------------------------------------------------------------------------------------------ 2. MMA issue rate
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;          // SWIZZLE_128B
    return d;
}
// a_mode: 0 = canonical A tile (8-row groups 1024 B apart), 1 = halo window (groups 1280 B apart, start shifted by 11 rows),
//         2 = wide halo window (groups 2304 B apart)
__global__ void __launch_bounds__(128) mma_probe(int N, int iters, int a_mode, int fp16, unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (48 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;     // 1.0 (fp16) / small (bf16)
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t fmt = fp16 ? 0u : ((1u << 7) | (1u << 10));
        const uint32_t idesc = (1u << 4) | fmt | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sa = smem_u32(smem) + (a_mode ? 11 * 128 : 0);
        const uint32_t sb = smem_u32(smem) + 48 * 1024;
        const uint32_t sbo_a = a_mode == 0 ? 1024 : (a_mode == 1 ? 1280 : 2304);
        const unsigned long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = make_desc(sa + k * 32, sbo_a);
                const uint64_t db = make_desc(sb + k * 32, 1024);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------ 3. SiLU arithmetic
__device__ __forceinline__ float silu_exp(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
__device__ __forceinline__ float silu_tanh(float v) {
    float h = 0.5f * v, t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
__device__ __forceinline__ __half2 silu_tanh_h2(__half2 v) {
    const __half2 h = __hmul2(v, __float2half2_rn(0.5f));
    uint32_t t, hi = *reinterpret_cast<const uint32_t*>(&h);
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(hi));
    return __hfma2(h, *reinterpret_cast<__half2*>(&t), h);
}
__device__ __forceinline__ __half2 silu_exp_h2(__half2 v) {        // x * 1 / (1 + 2^(-x log2 e)) in half2
    const __half2 e = h2exp2(__hmul2(v, __float2half2_rn(-1.4426950408889634f)));
    return __hmul2(v, h2rcp(__hadd2(e, __float2half2_rn(1.0f))));
}
template <int MODE>
__global__ void __launch_bounds__(256) silu_probe(int iters, float* out, float* err_out, unsigned long long* cycles) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = -6.0f + 12.0f * (float)((threadIdx.x * 8 + j) % 997) / 997.0f;
    float acc = 0.f;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            float a = x[j] + acc * 1e-30f, b = x[j + 1] - acc * 1e-30f;
            if (MODE == 0) { a = silu_exp(a); b = silu_exp(b); }
            else if (MODE == 1) { a = silu_tanh(a); b = silu_tanh(b); }
            else {
                __half2 h = __floats2half2_rn(a, b);
                h = MODE == 2 ? silu_tanh_h2(h) : silu_exp_h2(h);
                const float2 f = __half22float2(h); a = f.x; b = f.y;
            }
            acc += a + b;
        }
    }
    const unsigned long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    // accuracy: max |approx - exact| over a sweep, one block
    if (blockIdx.x == 0) {
        float worst = 0.f;
        for (int k = threadIdx.x; k < 65536; k += blockDim.x) {
            const float v = -12.0f + 24.0f * (float)k / 65536.0f;
            const double ex = (double)v / (1.0 + exp(-(double)v));
            float a;
            if (MODE == 0) a = silu_exp(v);
            else if (MODE == 1) a = silu_tanh(v);
            else { __half2 h = __floats2half2_rn(v, v); h = MODE == 2 ? silu_tanh_h2(h) : silu_exp_h2(h); a = __half22float2(h).x; }
            worst = fmaxf(worst, fabsf((float)((double)a - ex)));
        }
        err_out[threadIdx.x] = worst;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int sel = argc > 1 ? atoi(argv[1]) : 15;      // bit 0 TMA regions, 1 MMA, 2 SiLU, 3 TMA producers x box size
    int sms = 0, clk = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, max SM clock %d kHz\n", sms, clk);
    unsigned long long* d_cyc; CK(cudaMalloc(&d_cyc, 4096 * 8));
    std::vector<unsigned long long> cyc(4096);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    // ---- 1. TMA
    if (sel & 25) {
        void* fp = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
        EncodeTiledFn enc = (EncodeTiledFn)fp;
        const long long total_rows = (4LL << 30) / 128;       // 4 GB buffer
        void* buf; CK(cudaMalloc(&buf, (size_t)total_rows * 128)); CK(cudaMemset(buf, 0, (size_t)total_rows * 128));
        CUtensorMap map;
        cuuint64_t dims[2] = {64, (cuuint64_t)total_rows}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, kBoxRows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        CK(cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * 16384 + 1024));
        printf("\n== TMA 16 KB boxes, %d in flight per CTA: GB/s (event time), B/clk/SM (clock64)\n", kStages);
        struct Case { const char* name; int region_rows; int shared; } cases[] = {
            {"own 128 KB region (L2 hit)", 1024, 0}, {"own 256 KB region (L2 hit)", 2048, 0}, {"all CTAs same 1 MB (L2 hit)", 8192, 1},
            {"CTA pairs share 1 MB (L2 hit)", 8192, 2}, {"own 24 MB region (HBM stream)", 196608, 0}};
        const int grids[] = {37, 74, 148, 296, 444};
        for (auto& c : cases) for (int g : grids) {
            if (!(sel & 1)) break;
            const int iters = 4000;
            tma_probe<<<g, 128, kStages * 16384 + 1024>>>(map, 200, c.region_rows, c.shared, total_rows, d_cyc);      // warm L2
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            tma_probe<<<g, 128, kStages * 16384 + 1024>>>(map, iters, c.region_rows, c.shared, total_rows, d_cyc);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(cyc.data(), d_cyc, g * 8, cudaMemcpyDeviceToHost));
            double mean = 0; for (int i = 0; i < g; ++i) mean += (double)cyc[i]; mean /= g;
            const double bytes = (double)g * iters * 16384.0;
            printf("  %-34s grid %3d: %8.1f GB/s   %6.1f B/clk/CTA   %7.1f B/clk/SM\n", c.name, g, bytes / (ms * 1e-3) / 1e9, iters * 16384.0 / mean,
                   iters * 16384.0 / mean * ((g + sms - 1) / sms));
        }
        // ---- second experiment: producers per CTA x box size x ring depth, 1 CTA per SM, L2-resident regions
        if (sel & 8) {
        printf("\n== TMA, 1 CTA/SM (grid %d): producer warps x box bytes x boxes in flight per producer -> B/clk/SM, GB/s\n", sms);
        CK(cudaFuncSetAttribute(tma_probe_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int box_rows_l[] = {64, 128, 256};
        for (int br : box_rows_l) {
            CUtensorMap m2; cuuint32_t box2[2] = {64, (cuuint32_t)br};
            r = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            for (int prod = 1; prod <= 4; prod *= 2) for (int st = 2; st <= 8; st *= 2) {
                if ((size_t)prod * st * br * 128 > 190 * 1024) continue;
                const int iters = 3000;
                tma_probe_multi<<<sms, 128, 200 * 1024>>>(m2, 100, br, prod, st, total_rows, d_cyc); CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                tma_probe_multi<<<sms, 128, 200 * 1024>>>(m2, iters, br, prod, st, total_rows, d_cyc);
                CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                CK(cudaMemcpy(cyc.data(), d_cyc, sms * 8, cudaMemcpyDeviceToHost));
                double mean = 0; for (int i = 0; i < sms; ++i) mean += (double)cyc[i]; mean /= sms;
                printf("  producers %d  box %5d B  in flight %d: %7.1f B/clk/SM  %8.1f GB/s\n", prod, br * 128, st, (double)prod * iters * br * 128.0 / mean,
                       (double)sms * prod * iters * br * 128.0 / (ms * 1e-3) / 1e9);
            }
        }
        }
        if (sel & 16) {
            printf("\n== TMA 3-D boxes {64, rows, chunks}, 1 CTA/SM, 1 producer: box bytes x boxes in flight -> B/clk/SM, GB/s\n");
            CK(cudaFuncSetAttribute(tma_probe_3d, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            const int chunks_l[] = {1, 2, 3, 4, 6};
            const int rows_l[] = {128, 256};
            for (int rows : rows_l) for (int ch : chunks_l) {
                const int box_bytes = rows * ch * 128;
                if (box_bytes > 96 * 1024) continue;
                // memory viewed as [row][8 chunks][64]: row pitch 1024 B
                CUtensorMap m3; cuuint64_t d3[3] = {64, (cuuint64_t)(total_rows / 8), 8}; cuuint64_t s3[2] = {1024, 128};
                cuuint32_t b3[3] = {64, (cuuint32_t)rows, (cuuint32_t)ch}; cuuint32_t e3[3] = {1, 1, 1};
                r = enc(&m3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encode 3d failed %d (rows %d chunks %d)\n", (int)r, rows, ch); continue; }
                for (int st = 1; st <= 4; ++st) {
                    if ((size_t)st * box_bytes > 190 * 1024) continue;
                    const int iters = 2000, region_rows = 1024;      // 1 MB per CTA (rows of 1 KB): 148 MB > L2?  keep 512 rows = 512 KB
                    tma_probe_3d<<<sms, 128, 200 * 1024>>>(m3, 50, box_bytes, rows, st, 512, d_cyc); CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0));
                    tma_probe_3d<<<sms, 128, 200 * 1024>>>(m3, iters, box_bytes, rows, st, 512, d_cyc);
                    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    CK(cudaMemcpy(cyc.data(), d_cyc, sms * 8, cudaMemcpyDeviceToHost));
                    double mean = 0; for (int i = 0; i < sms; ++i) mean += (double)cyc[i]; mean /= sms;
                    printf("  box %6d B (%3d rows x %d chunks) in flight %d: %7.1f B/clk/SM  %8.1f GB/s   %6.0f clk/box\n", box_bytes, rows, ch, st,
                           (double)iters * box_bytes / mean, (double)sms * iters * box_bytes / (ms * 1e-3) / 1e9, mean / iters);
                    (void)region_rows;
                }
            }
        }
        CK(cudaFree(buf));
    }
    // ---- 2. MMA
    if (sel & 2) {
        CK(cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 84 * 1024));
        printf("\n== tcgen05.mma kind::f16 M=128 K=16 from shared memory: clk per MMA per CTA, TFLOP/s chip (event time)\n");
        const int Ns[] = {64, 128, 256};
        for (int fp16 = 0; fp16 < 2; ++fp16) for (int a_mode = 0; a_mode < 3; ++a_mode) for (int N : Ns) for (int per_sm = 1; per_sm <= 2; ++per_sm) {
            if (fp16 && a_mode) continue;
            const int iters = 20000, g = sms * per_sm;
            mma_probe<<<g, 128, 84 * 1024>>>(N, 100, a_mode, fp16, d_cyc); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            mma_probe<<<g, 128, 84 * 1024>>>(N, iters, a_mode, fp16, d_cyc);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(cyc.data(), d_cyc, g * 8, cudaMemcpyDeviceToHost));
            double mean = 0; for (int i = 0; i < g; ++i) mean += (double)cyc[i]; mean /= g;
            const double flops = (double)g * iters * 4 * 2.0 * 128 * N * 16;
            printf("  %s a_mode %d N %3d, %d CTA/SM: %7.1f clk/MMA/CTA  %8.1f TFLOP/s\n", fp16 ? "fp16" : "bf16", a_mode, N, per_sm, mean / (iters * 4.0), flops / (ms * 1e-3) / 1e12);
        }
    }
    // ---- 3. SiLU
    if (sel & 4) {
        float *d_out, *d_err; CK(cudaMalloc(&d_out, 148 * 8 * 256 * 4)); CK(cudaMalloc(&d_err, 256 * 4));
        std::vector<float> err(256);
        printf("\n== SiLU variants: elements / clk / SM (8 CTAs x 256 threads per SM), max |error| on [-12, 12]\n");
        const char* names[] = {"ex2 + rcp (f32, 2 MUFU)", "tanh.approx.f32 (1 MUFU)", "tanh.approx.f16x2", "ex2.f16x2 + rcp.f16x2"};
        for (int mode = 0; mode < 4; ++mode) {
            const int iters = 4000, g = sms * 8;
            auto launch = [&](int it) {
                if (mode == 0) silu_probe<0><<<g, 256>>>(it, d_out, d_err, d_cyc);
                else if (mode == 1) silu_probe<1><<<g, 256>>>(it, d_out, d_err, d_cyc);
                else if (mode == 2) silu_probe<2><<<g, 256>>>(it, d_out, d_err, d_cyc);
                else silu_probe<3><<<g, 256>>>(it, d_out, d_err, d_cyc);
            };
            launch(10); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); launch(iters); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(cyc.data(), d_cyc, g * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(err.data(), d_err, 256 * 4, cudaMemcpyDeviceToHost));
            double mean = 0; for (int i = 0; i < g; ++i) mean += (double)cyc[i]; mean /= g;
            float worst = 0; for (float v : err) worst = v > worst ? v : worst;
            const double elems_per_sm = 8.0 * 256 * 8 * iters;
            printf("  %-28s %6.1f elem/clk/SM (clock64)  %8.1f Gelem/s chip   max abs err %.3e\n", names[mode], elems_per_sm / mean,
                   (double)g * 256 * 8 * iters / (ms * 1e-3) / 1e9, worst);
        }
    }
    return 0;
}
