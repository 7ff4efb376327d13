// b2t_conv.cu -- conv2d + bias + SiLU as an implicit GEMM on the 5th-gen tensor cores (sm_100a).
//
// Replaces ``Conv.fuseforward`` (models/common.py:110-111: act(conv(x)) with BN folded by
// utils/torch_utils.py:181-201) and the linear 1x1 head convs of ``Detect`` (models/yolo.py:44)
// for the YOLOv7-w6 / tiny graphs (k in {1,3}, s in {1,2}, pad k//2, groups 1).
//
//   D[pixel, cout] = sum_{kh,kw,cin} X[n, ho*s+kh-p, wo*s+kw-p, cin] * W[cout, kh, kw, cin]
//
// Layout: activations NHWC bf16 (a tensor may be a channel slice of a wider concat buffer: pitch !=
// C), weights [Cout][KH][KW][Cin] bf16 (K-major), bias fp32, accumulation fp32 in TMEM.
//
// One CTA computes a 128-pixel x BLOCK_N-channel output tile:
//   * the 128 pixels are a TH x TW spatial patch of one image (TH*TW = 128).  For filter tap
//     (kh,kw) and channel chunk kc the A operand is ONE TMA box {BK ch, TW, TH, 1} of the 4-D
//     tensor map (C, W, H, N) at coordinates (kc*BK, wo0*s+kw-p, ho0*s+kh-p, n): out-of-bounds
//     rows/cols are zero-filled by the TMA unit (that IS the padding), stride-2 convs use the tensor
//     map's element strides, and the box lands in shared memory as 128 rows of BK*2 bytes with the
//     128-/64-/32-byte swizzle -- exactly the canonical K-major UMMA operand layout.  No im2col.
//   * B operand: 2-D map (K, Cout), box {BK, BLOCK_N}.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread, tcgen05.mma
//     cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16 per instruction), warps 2..5 = epilogue:
//     tcgen05.ld 32 lanes x 32 columns -> +bias -> SiLU -> bf16 (or fp32) -> 16-byte global stores
//     straight into the consumer's concat buffer (concat-by-address).
//   * STAGES-deep shared-memory ring with full/empty mbarriers; tcgen05.commit releases a stage.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <cstdio>
#include "../../include/b200track.h"

namespace {

constexpr int kMaxStages = 6;
constexpr int kThreads = 192;      // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kTileM = 128;

struct ConvParams {
    int N, H, W, Cin;              // input geometry (Cin = channels of the slice read)
    int Ho, Wo, Cout;              // output geometry
    int KH, KW, stride, pad;
    int TH, TW;                    // spatial tile, TH*TW == 128
    int BK;                        // K chunk: 64 (SW128), 32 (SW64) or 16 (SW32) channels
    int BN;                        // output channels per CTA, multiple of 16, <= 256
    int tiles_w, tiles_h;          // tiles per image
    int out_pitch;                 // elements per output pixel (concat buffer width)
    int out_coff;                  // channel offset inside the output buffer
    int act;                       // 1 = SiLU, 0 = linear
    int out_f32;                   // 1 = fp32 output (head), 0 = bf16
    int flat;                      // 1 = 1x1/s1: pixels are the flattened N*H*W axis (2-D A map), TH/TW unused
    int stages;                    // shared-memory ring depth (<= kMaxStages)
    int tmem_cols;                 // power of two >= BN (>= 32)
    long long total_pix;           // N*Ho*Wo (flat mode bound)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand tile: rows of (BK*2) bytes, 8-row swizzle atoms stacked every 8*(BK*2) bytes.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, int row_bytes) {
    const uint32_t sbo = (uint32_t)(8 * row_bytes) >> 4;
    const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);     // SWIZZLE_128B / 64B / 32B
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major), 16-byte units
    d |= (uint64_t)(sbo & 0x3fff) << 32;    // stride byte offset
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;
    return d;
}

__device__ __forceinline__ float silu(float v) { return v / (1.0f + __expf(-v)); }

__global__ void __launch_bounds__(kThreads)
conv_bias_act_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const float* __restrict__ bias, void* __restrict__ out, const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_bytes = kTileM * p.BK * 2;
    const int b_bytes = p.BN * p.BK * 2;
    const int stage_bytes = ((a_bytes + b_bytes + 1023) / 1024) * 1024;
    // swizzled operand tiles need 1024-byte alignment in the shared window (slack is reserved by the host)
    uint8_t* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    const int kStages = p.stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + kStages * stage_bytes);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    // ---- tile coordinates
    const int n0 = blockIdx.y * p.BN;
    int img = 0, ho0 = 0, wo0 = 0;
    long long pix0 = 0;
    if (p.flat) {
        pix0 = (long long)blockIdx.x * kTileM;
    } else {
        const int per_img = p.tiles_w * p.tiles_h;
        img = blockIdx.x / per_img;
        const int t = blockIdx.x % per_img;
        ho0 = (t / p.tiles_w) * p.TH;
        wo0 = (t % p.tiles_w) * p.TW;
    }
    const int kchunks = p.Cin / p.BK;
    const int ktotal = p.KH * p.KW * kchunks;

    // ---- one-time setup
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kt = 0; kt < ktotal; ++kt) {
                const int tap = kt / kchunks, kc = kt % kchunks;
                const int kh = tap / p.KW, kw = tap % p.KW;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = tiles + stage * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                mbar_expect_tx(&full_bar[stage], (uint32_t)(a_bytes + b_bytes));
                if (p.flat) tma_load_2d(sa, &map_a, &full_bar[stage], kc * p.BK, (int)pix0);
                else tma_load_4d(sa, &map_a, &full_bar[stage], kc * p.BK, wo0 * p.stride + kw - p.pad, ho0 * p.stride + kh - p.pad, img);
                tma_load_2d(sb, &map_b, &full_bar[stage], tap * p.Cin + kc * p.BK, n0);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
            const int row_bytes = p.BK * 2;
            int stage = 0; uint32_t phase = 0;
            for (int kt = 0; kt < ktotal; ++kt) {
                mbar_wait(&full_bar[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(tiles + stage * stage_bytes);
                const uint32_t sb = sa + a_bytes;
                for (int k = 0; k < p.BK / 16; ++k) {
                    const uint64_t da = make_smem_desc(sa + k * 32, row_bytes);
                    const uint64_t db = make_smem_desc(sb + k * 32, row_bytes);
                    umma_bf16(tmem_base, da, db, idesc, (kt | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);          // stage reusable once these MMAs retire
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tmem_full);                       // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31
        const int q = warp & 3;
        const int row = q * 32 + lane;                    // pixel row inside the tile
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        long long pix;
        bool valid;
        if (p.flat) { pix = pix0 + row; valid = pix < p.total_pix; }
        else {
            const int ho = ho0 + row / p.TW, wo = wo0 + row % p.TW;
            valid = ho < p.Ho && wo < p.Wo;
            pix = ((long long)img * p.Ho + ho) * p.Wo + wo;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int cbase = n0 + c0;
            if (valid && cbase < p.Cout) {
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int c = cbase + i;
                    float x = __uint_as_float(v[i]) + (c < p.Cout ? __ldg(bias + c) : 0.f);
                    f[i] = p.act ? silu(x) : x;
                }
                const int nvalid = p.Cout - cbase < 32 ? p.Cout - cbase : 32;
                if (p.out_f32) {
                    float* o = reinterpret_cast<float*>(out) + pix * p.out_pitch + p.out_coff + cbase;
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
                    } else {
                        for (int i = 0; i < nvalid; ++i) o[i] = f[i];
                    }
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + pix * p.out_pitch + p.out_coff + cbase;
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(f[i], f[i + 1]), h1 = __floats2bfloat162_rn(f[i + 2], f[i + 3]);
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[i + 4], f[i + 5]), h3 = __floats2bfloat162_rn(f[i + 6], f[i + 7]);
                            uint4 u;
                            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
                            *reinterpret_cast<uint4*>(o + i) = u;
                        }
                    } else {
                        for (int i = 0; i < nvalid; ++i) o[i] = __float2bfloat16_rn(f[i]);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
thread_local std::string g_conv_err;
int cfail(int code, const std::string& m) { g_conv_err = m; return code; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
CUtensorMapSwizzle swizzle_for(int bk) {
    return bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

}  // namespace

struct b2t_conv_plan {
    CUtensorMap map_a, map_b;
    ConvParams p;
    const float* bias;
    void* out;
    dim3 grid;
    size_t smem;
};

extern "C" const char* b2t_conv_last_error(void) { return g_conv_err.c_str(); }

extern "C" int b2t_conv_plan_create(const b2t_conv_desc* d, b2t_conv_plan** out_plan) {
    if (!d || !out_plan) return cfail(B2T_EINVAL, "b2t_conv_plan_create: null argument");
    if (!(d->kh == d->kw && (d->kh == 1 || d->kh == 3)) || !(d->stride == 1 || d->stride == 2))
        return cfail(B2T_EINVAL, "b2t_conv_plan_create: only k in {1,3}, stride in {1,2}");
    int bk = d->cin % 64 == 0 ? 64 : (d->cin % 32 == 0 ? 32 : (d->cin % 16 == 0 ? 16 : 0));
    if (!bk) return cfail(B2T_EINVAL, "b2t_conv_plan_create: Cin must be a multiple of 16");
    if (d->in_pitch % 8 || d->in_coff % 8 || d->out_coff % 8)
        return cfail(B2T_EINVAL, "b2t_conv_plan_create: pitches / offsets must keep 16-byte alignment");
    EncodeTiledFn enc = get_encode();
    if (!enc) return cfail(B2T_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    b2t_conv_plan* pl = new b2t_conv_plan();
    ConvParams& p = pl->p;
    p.N = d->n; p.H = d->h; p.W = d->w; p.Cin = d->cin; p.Cout = d->cout;
    p.KH = d->kh; p.KW = d->kw; p.stride = d->stride; p.pad = d->kh / 2;
    p.Ho = (d->h + 2 * p.pad - d->kh) / d->stride + 1;
    p.Wo = (d->w + 2 * p.pad - d->kw) / d->stride + 1;
    p.BK = bk;
    const int cout_pad = (d->cout + 15) / 16 * 16;
    int bn = cout_pad;
    if (bn > 256) { bn = 256; for (int cand = 256; cand >= 64; cand -= 16) if (cout_pad % cand == 0) { bn = cand; break; } }
    if (d->block_n > 0) bn = d->block_n;
    if (bn % 16 || bn > 256 || bn < 16) return cfail(B2T_EINVAL, "b2t_conv_plan_create: bad BLOCK_N");
    p.BN = bn;
    p.out_pitch = d->out_pitch; p.out_coff = d->out_coff; p.act = d->act; p.out_f32 = d->out_f32;
    p.flat = (d->kh == 1 && d->stride == 1) ? 1 : 0;
    p.total_pix = (long long)p.N * p.Ho * p.Wo;
    if (p.flat) { p.TH = 1; p.TW = 128; p.tiles_w = p.tiles_h = 0; }
    else {
        int tw = 16;
        if (p.Wo % 16 != 0) { tw = (p.Wo % 8 == 0) ? 8 : 4; }
        if (d->tile_w > 0) tw = d->tile_w;
        p.TW = tw; p.TH = 128 / tw;
        p.tiles_w = (p.Wo + p.TW - 1) / p.TW; p.tiles_h = (p.Ho + p.TH - 1) / p.TH;
    }
    // ---- tensor maps
    const CUtensorMapSwizzle sw = swizzle_for(bk);
    char* a_base = reinterpret_cast<char*>(const_cast<void*>(d->x)) + (size_t)d->in_coff * 2;
    CUresult r;
    if (p.flat) {
        cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)p.total_pix};
        cuuint64_t strides[1] = {(cuuint64_t)d->in_pitch * 2};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)kTileM};
        cuuint32_t es[2] = {1, 1};
        r = enc(&pl->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a_base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)d->in_pitch * 2, (cuuint64_t)d->in_pitch * 2 * p.W, (cuuint64_t)d->in_pitch * 2 * p.W * p.H};
        // with element strides the box extent is given in INPUT elements: TW outputs at stride s span TW*s inputs
        cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(p.TW * p.stride), (cuuint32_t)(p.TH * p.stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)p.stride, (cuuint32_t)p.stride, 1};
        r = enc(&pl->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a_base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { delete pl; return cfail(B2T_ECUDA, "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r)); }
    {
        const cuuint64_t K = (cuuint64_t)p.KH * p.KW * p.Cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)d->cout_rows};
        cuuint64_t strides[1] = {K * 2};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)bn};
        cuuint32_t es[2] = {1, 1};
        r = enc(&pl->map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w_packed), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete pl; return cfail(B2T_ECUDA, "cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r)); }
    }
    pl->bias = d->bias; pl->out = d->y;
    const int a_bytes = kTileM * bk * 2, b_bytes = bn * bk * 2;
    const int stage_bytes = ((a_bytes + b_bytes + 1023) / 1024) * 1024;
    const int ktotal = p.KH * p.KW * (p.Cin / bk);
    // ring depth: no deeper than the K loop, and shallow enough that 2+ CTAs share an SM (one CTA's
    // epilogue then overlaps another's MMAs -- the kernel itself is not persistent)
    int stages = ktotal < 4 ? ktotal : 4;
    if (d->stages > 0) stages = d->stages < kMaxStages ? d->stages : kMaxStages;
    if (stages > ktotal) stages = ktotal;
    while (stages > 2 && (size_t)stages * stage_bytes > 100 * 1024) --stages;
    p.stages = stages;
    int tc = 32; while (tc < bn) tc <<= 1;
    p.tmem_cols = tc;
    pl->smem = (size_t)stages * stage_bytes + 256 + 1024;
    const int tiles_m = p.flat ? (int)((p.total_pix + kTileM - 1) / kTileM) : p.N * p.tiles_w * p.tiles_h;
    pl->grid = dim3(tiles_m, (cout_pad + bn - 1) / bn, 1);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_bias_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            delete pl; return cfail(B2T_ECUDA, "cannot raise dynamic shared memory for conv kernel");
        }
        attr_set = true;
    }
    *out_plan = pl;
    return B2T_OK;
}

extern "C" void b2t_conv_plan_destroy(b2t_conv_plan* pl) { delete pl; }

extern "C" double b2t_conv_plan_flops(const b2t_conv_plan* pl) {
    const ConvParams& p = pl->p;
    return 2.0 * (double)p.total_pix * p.Cout * p.KH * p.KW * p.Cin;
}

extern "C" int b2t_conv_run(const b2t_conv_plan* pl, void* stream) {
    if (!pl) return cfail(B2T_EINVAL, "b2t_conv_run: null plan");
    conv_bias_act_kernel<<<pl->grid, kThreads, pl->smem, (cudaStream_t)stream>>>(pl->map_a, pl->map_b, pl->bias, pl->out, pl->p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cfail(B2T_ECUDA, std::string("conv launch: ") + cudaGetErrorString(e));
    return B2T_OK;
}
