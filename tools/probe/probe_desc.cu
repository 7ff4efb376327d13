// probe_desc.cu -- how does a tcgen05 shared-memory descriptor (SWIZZLE_128B, K-major) behave when the matrix start
// address is shifted by whole 128-byte rows and the 8-row groups are strided by a non-multiple of 1024 bytes?
// (needed for a "halo tile" 3x3 convolution: one (TH+2) x (TW+2) input tile in shared memory, nine shifted A windows.)
// A is filled with a known pattern directly in the TMA 128B-swizzle layout, B = identity, so D[m][n] = A_row(m)[n].
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;          // SWIZZLE_128B
    return d;
}

// mode: 0 value = row index, 1 value = channel index
__global__ void probe(int shift_rows, int sbo_bytes, int use_base_off, int mode, float* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int R = 256;                                  // pixel rows in the halo buffer
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);               // R x 64 bf16 = 32 KB
    uint8_t* Bm = smem + R * 128;                                            // 64 x 64 identity, swizzled, 8 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < R * 64; i += blockDim.x) {
        const int r = i >> 6, c = i & 63, j = c >> 3;
        const float v = mode == 0 ? (float)r : (float)c;
        *reinterpret_cast<__nv_bfloat16*>(smem + r * 128 + ((j ^ (r & 7)) << 4) + (c & 7) * 2) = __float2bfloat16(v);
    }
    for (int i = tid; i < 64 * 64; i += blockDim.x) {
        const int n = i >> 6, k = i & 63, j = k >> 3;
        *reinterpret_cast<__nv_bfloat16*>(Bm + n * 128 + ((j ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(n == k ? 1.f : 0.f);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sa = smem_u32(A) + (uint32_t)shift_rows * 128u;
        const uint32_t sb = smem_u32(Bm);
        for (int k = 0; k < 4; ++k) {
            const uint32_t a_addr = sa + k * 32;
            const uint64_t da = make_desc(a_addr, (uint32_t)sbo_bytes, use_base_off ? ((a_addr >> 7) & 7) : 0);
            const uint64_t db = make_desc(sb + k * 32, 1024, 0);
            const uint32_t acc = k != 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // all 4 warps wait, then read their 32 lanes
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    float* d; cudaMalloc(&d, 128 * 64 * 4);
    std::vector<float> rows(128 * 64), chans(128 * 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int sbos[] = {1024, 1152, 1280};
    const int shifts[] = {0, 1, 2, 7, 10, 11, 21};
    for (int sbo : sbos) for (int shift : shifts) for (int bo = 0; bo < 2; ++bo) {
        probe<<<1, 128, 48 * 1024>>>(shift, sbo, bo, 0, d); cudaMemcpy(rows.data(), d, rows.size() * 4, cudaMemcpyDeviceToHost);
        probe<<<1, 128, 48 * 1024>>>(shift, sbo, bo, 1, d); cudaMemcpy(chans.data(), d, chans.size() * 4, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d shift %d base_off %d: CUDA error %s\n", sbo, shift, bo, cudaGetErrorString(e)); return 1; }
        int bad_row = 0, bad_ch = 0, row_uniform = 0;
        for (int m = 0; m < 128; ++m) {
            const int want = shift + (m / 8) * (sbo / 128) + (m % 8);
            bool uni = true;
            for (int n = 0; n < 64; ++n) {
                if (rows[m * 64 + n] != (float)want) ++bad_row;
                if (chans[m * 64 + n] != (float)n) ++bad_ch;
                if (rows[m * 64 + n] != rows[m * 64]) uni = false;
            }
            row_uniform += uni;
        }
        printf("sbo %4d shift %2d base_off %d: wrong-row elems %4d  wrong-channel elems %4d  rows-uniform %3d/128 | m=0..9 row:", sbo, shift, bo, bad_row, bad_ch, row_uniform);
        for (int m = 0; m < 10; ++m) printf(" %g", rows[m * 64]);
        printf(" | m=1 chan chunks:");
        for (int j = 0; j < 8; ++j) printf(" %g", chans[1 * 64 + j * 8]);
        printf("\n");
    }
    return 0;
}
