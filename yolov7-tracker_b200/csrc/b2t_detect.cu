// b2t_detect.cu -- the memory-bound glue around the convolutions of the YOLOv7 detector branch, and NMS.
//
//   image_reorg_kernel     ReOrg (models/common.py:48-53) + NCHW fp32 -> NHWC bf16 (+ zero pad to 16 ch)
//   upsample2x_kernel      nn.Upsample(None, 2, 'nearest'), written straight into the concat buffer
//   spp_pool_kernel        the three stride-1 max-pools (5/9/13) of SPPCSPC (models/common.py:271,278)
//   detect_decode_kernel   Detect.forward inference branch (models/yolo.py:44-55): sigmoid, grid / anchor decode
//   nms_*                  utils/general.py:607-695 non_max_suppression (best-class path) incl. the
//                          torchvision.ops.nms greedy suppression (:679) and the 300-detection cap (:680-681),
//                          + scale_coords / clip / round of tracker/track.py:239-240.
// All HBM-bound elementwise / scan work: coalesced along channels, 16-byte accesses where the layout allows.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include "../../include/b200track.h"

namespace {
thread_local std::string g_det_err;
int dfail(int code, const char* m) { g_det_err = m; return code; }
int dcheck(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_det_err = std::string(what) + ": " + cudaGetErrorString(e); return B2T_ECUDA; }
    return B2T_OK;
}

// ---------------------------------------------------------------- ReOrg + layout change
// out[b][y][x][phase*3 + c] = img[b][c][2y + dy][2x + dx], phase order (dy,dx) = (0,0),(1,0),(0,1),(1,1)
__global__ void image_reorg_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H, int W) {
    const int H2 = H / 2, W2 = W / 2;
    const long long total = (long long)B * H2 * W2;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(p % W2), y = (int)((p / W2) % H2), b = (int)(p / ((long long)W2 * H2));
        __nv_bfloat16 v[16];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const int dy = ph & 1, dx = ph >> 1;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v[ph * 3 + c] = __float2bfloat16_rn(img[(((long long)b * 3 + c) * H + 2 * y + dy) * W + 2 * x + dx]);
        }
        v[12] = v[13] = v[14] = v[15] = __float2bfloat16_rn(0.f);
        uint4* o = reinterpret_cast<uint4*>(out + p * 16);
        o[0] = *reinterpret_cast<uint4*>(&v[0]);
        o[1] = *reinterpret_cast<uint4*>(&v[8]);
    }
}

// ---------------------------------------------------------------- nearest x2 upsample, 8 channels (16 B) per thread
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ src, int sp, int sc, __nv_bfloat16* __restrict__ dst, int dp, int dc,
                                  int B, int H, int W, int C) {
    const int cv = C / 8;
    const long long total = (long long)B * (2 * H) * (2 * W) * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long p = i / cv;
        const int x = (int)(p % (2 * W)), y = (int)((p / (2 * W)) % (2 * H)), b = (int)(p / ((long long)4 * W * H));
        const uint4 v = *reinterpret_cast<const uint4*>(src + (((long long)b * H + y / 2) * W + x / 2) * sp + sc + c8 * 8);
        *reinterpret_cast<uint4*>(dst + p * dp + dc + c8 * 8) = v;
    }
}

// ---------------------------------------------------------------- SPP max-pools 5 / 9 / 13 (stride 1, pad k/2)
__device__ __forceinline__ __nv_bfloat162 bmax2(__nv_bfloat162 a, __nv_bfloat162 b) { return __hmax2(a, b); }
__global__ void spp_pool_kernel(__nv_bfloat16* __restrict__ buf, int pitch, int C, int B, int H, int W) {
    const int cv = C / 2;
    const long long total = (long long)B * H * W * cv;
    const __nv_bfloat162 ninf = __floats2bfloat162_rn(-3.0e38f, -3.0e38f);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c2 = (int)(i % cv);
        long long p = i / cv;
        const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
        __nv_bfloat162 m5 = ninf, m9 = ninf, m13 = ninf;
        for (int dy = -6; dy <= 6; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -6; dx <= 6; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(buf + (((long long)b * H + yy) * W + xx) * pitch + c2 * 2);
                m13 = bmax2(m13, v);
                if (dy >= -4 && dy <= 4 && dx >= -4 && dx <= 4) m9 = bmax2(m9, v);
                if (dy >= -2 && dy <= 2 && dx >= -2 && dx <= 2) m5 = bmax2(m5, v);
            }
        }
        __nv_bfloat16* o = buf + p * pitch + c2 * 2;
        *reinterpret_cast<__nv_bfloat162*>(o + C) = m5;
        *reinterpret_cast<__nv_bfloat162*>(o + 2 * C) = m9;
        *reinterpret_cast<__nv_bfloat162*>(o + 3 * C) = m13;
    }
}

// ---------------------------------------------------------------- Detect decode
// raw [B][H][W][rp] fp32 (channel = a*no + o)  ->  pred [B][Ntot][no] rows level_off + (a*H + y)*W + x
__global__ void detect_decode_kernel(const float* __restrict__ raw, int rp, float* __restrict__ pred, int B, int H, int W, int na, int no,
                                     long long level_off, long long Ntot, float stride, float a0w, float a0h, float a1w, float a1h,
                                     float a2w, float a2h) {
    const long long total = (long long)B * H * W * na * no;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(i % no);
        long long t = i / no;
        const int x = (int)(t % W); t /= W;
        const int y = (int)(t % H); t /= H;
        const int a = (int)(t % na);
        const int b = (int)(t / na);
        const float r = raw[(((long long)b * H + y) * W + x) * rp + a * no + o];
        float s = 1.0f / (1.0f + expf(-r));
        if (o == 0) s = (s * 2.0f - 0.5f + (float)x) * stride;
        else if (o == 1) s = (s * 2.0f - 0.5f + (float)y) * stride;
        else if (o == 2) { const float aw = a == 0 ? a0w : (a == 1 ? a1w : a2w); s = (s * 2.0f) * (s * 2.0f) * aw; }
        else if (o == 3) { const float ah = a == 0 ? a0h : (a == 1 ? a1h : a2h); s = (s * 2.0f) * (s * 2.0f) * ah; }
        pred[((long long)b * Ntot + level_off + ((long long)a * H + y) * W + x) * no + o] = s;
    }
}

// ---------------------------------------------------------------- NMS
// candidate record: x1 y1 x2 y2 conf cls (float) + original row index
struct Cand { float x1, y1, x2, y2, conf, cls; int idx; int pad; };

__global__ void nms_filter_kernel(const float* __restrict__ pred, int N, int no, float conf_thres, Cand* __restrict__ cand, int* __restrict__ count, int maxc) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* r = pred + ((long long)b * N + i) * no;
    const float obj = r[4];
    if (!(obj > conf_thres)) return;                                 // xc = prediction[..., 4] > conf_thres
    float best = -1.f; int bj = 0;
    for (int c = 5; c < no; ++c) { const float v = r[c] * obj; if (v > best) { best = v; bj = c - 5; } }   // x[:, 5:] *= x[:, 4:5]; max(1)
    if (!(best > conf_thres)) return;
    const int slot = atomicAdd(&count[b], 1);
    if (slot >= maxc) return;
    Cand cd;
    cd.x1 = r[0] - r[2] / 2; cd.y1 = r[1] - r[3] / 2; cd.x2 = r[0] + r[2] / 2; cd.y2 = r[1] + r[3] / 2;   // xywh2xyxy
    cd.conf = best; cd.cls = (float)bj; cd.idx = i; cd.pad = 0;
    cand[(long long)b * maxc + slot] = cd;
}

// rank by (conf desc, original index asc): a counting sort key, deterministic whatever the atomics' order was
__global__ void nms_rank_kernel(const Cand* __restrict__ cand, const int* __restrict__ count, int maxc, int* __restrict__ rank) {
    __shared__ float sc[256];
    __shared__ int si[256];
    const int b = blockIdx.y;
    const int n = min(count[b], maxc);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)(blockIdx.x * blockDim.x) >= n) return;
    const Cand* cb = cand + (long long)b * maxc;
    float my = 0.f; int myi = 0;
    if (i < n) { my = cb[i].conf; myi = cb[i].idx; }
    int r = 0;
    for (int j0 = 0; j0 < n; j0 += 256) {
        const int j = j0 + threadIdx.x;
        __syncthreads();
        if (j < n) { sc[threadIdx.x] = cb[j].conf; si[threadIdx.x] = cb[j].idx; }
        __syncthreads();
        const int lim = min(256, n - j0);
        if (i < n) for (int k = 0; k < lim; ++k) r += (sc[k] > my || (sc[k] == my && si[k] < myi)) ? 1 : 0;
    }
    if (i < n) rank[(long long)b * maxc + i] = r;
}

__global__ void nms_scatter_kernel(const Cand* __restrict__ cand, const int* __restrict__ count, const int* __restrict__ rank, int maxc, int max_nms,
                                   float max_wh, float4* __restrict__ sbox, Cand* __restrict__ sorted) {
    const int b = blockIdx.y;
    const int n = min(count[b], maxc);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = rank[(long long)b * maxc + i];
    if (r >= max_nms) return;                                        // keep the max_nms best (utils/general.py:673-674)
    const Cand cd = cand[(long long)b * maxc + i];
    const float off = cd.cls * max_wh;                               // class offset (:677-678)
    sbox[(long long)b * max_nms + r] = make_float4(cd.x1 + off, cd.y1 + off, cd.x2 + off, cd.y2 + off);
    sorted[(long long)b * max_nms + r] = cd;
}

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {     // torchvision nms_kernel devIoU
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    const float inter = w * h;
    const float sa = (a.z - a.x) * (a.w - a.y), sb = (b.z - b.x) * (b.w - b.y);
    return inter / (sa + sb - inter) > thr;
}

// 64 x 64 tiles of the suppression matrix, one 64-bit word per (row, column block).  The number of
// candidates is only known on the device, so a fixed grid strides over the upper-triangular tiles.
__global__ void nms_mask_kernel(const float4* __restrict__ sbox, const int* __restrict__ count, int maxc, int max_nms, float thr,
                                unsigned long long* __restrict__ mask, int words) {
    const int b = blockIdx.y;
    const int n = min(min(count[b], maxc), max_nms);
    const int nw = (n + 63) / 64;
    __shared__ float4 cols[64];
    const float4* sb = sbox + (long long)b * max_nms;
    for (long long t = blockIdx.x; t < (long long)nw * nw; t += gridDim.x) {
        const int rb = (int)(t / nw), cbk = (int)(t % nw);
        if (cbk < rb) continue;
        __syncthreads();
        const int cj = cbk * 64 + threadIdx.x;
        if (cj < n) cols[threadIdx.x] = sb[cj];
        __syncthreads();
        const int i = rb * 64 + threadIdx.x;
        if (i >= n) continue;
        const float4 me = sb[i];
        unsigned long long bits = 0;
        const int lim = min(64, n - cbk * 64);
        for (int k = (rb == cbk ? threadIdx.x + 1 : 0); k < lim; ++k)
            if (iou_gt(me, cols[k], thr)) bits |= 1ull << k;
        mask[((long long)b * max_nms + i) * words + cbk] = bits;
    }
}

// greedy scan in score order; one CTA per image, 64 candidates per step: the diagonal 64 x 64 block is
// resolved by one thread with bit operations on words staged in shared memory, then all threads OR the
// kept rows' masks into the removal bitmap of the later blocks.  Stops at max_det kept rows.
// Writes rows [x1 y1 x2 y2 conf cls] after scale_coords (gain / pad) + clip + round when post != 0.
__global__ void nms_select_kernel(const Cand* __restrict__ sorted, const unsigned long long* __restrict__ mask, const int* __restrict__ count,
                                  int maxc, int max_nms, int words, int max_det, float* __restrict__ out, int* __restrict__ out_count,
                                  int post, float gain, float padw, float padh, float img_w, float img_h) {
    extern __shared__ unsigned long long remv[];
    __shared__ unsigned long long diag[64];
    __shared__ int kept_rows[64];
    __shared__ int s_nk, s_keep;
    const int b = blockIdx.x;
    const int n = min(min(count[b], maxc), max_nms);
    const int nw = (n + 63) / 64;
    for (int w = threadIdx.x; w < nw; w += blockDim.x) remv[w] = 0;
    if (threadIdx.x == 0) s_keep = 0;
    __syncthreads();
    const Cand* sc = sorted + (long long)b * max_nms;
    const unsigned long long* mb = mask + (long long)b * max_nms * words;
    for (int blk = 0; blk < nw; ++blk) {
        if (s_keep >= max_det) break;
        const int i0 = blk * 64;
        if (threadIdx.x < 64) { const int i = i0 + threadIdx.x; diag[threadIdx.x] = i < n ? mb[(long long)i * words + blk] : 0ull; }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cur = remv[blk];
            int nk = 0, k = s_keep;
            const int lim = min(64, n - i0);
            for (int q = 0; q < lim && k < max_det; ++q) {
                if ((cur >> q) & 1ull) continue;
                kept_rows[nk++] = i0 + q;
                cur |= diag[q];
                ++k;
            }
            s_nk = nk;
        }
        __syncthreads();
        const int nk = s_nk, k0 = s_keep;
        for (int q = threadIdx.x; q < nk; q += blockDim.x) {           // emit the kept rows
            const Cand cd = sc[kept_rows[q]];
            float x1 = cd.x1, y1 = cd.y1, x2 = cd.x2, y2 = cd.y2;
            if (post) {
                x1 = (x1 - padw) / gain; x2 = (x2 - padw) / gain; y1 = (y1 - padh) / gain; y2 = (y2 - padh) / gain;   // scale_coords
                x1 = fminf(fmaxf(x1, 0.f), img_w); x2 = fminf(fmaxf(x2, 0.f), img_w);                                  // clip_coords
                y1 = fminf(fmaxf(y1, 0.f), img_h); y2 = fminf(fmaxf(y2, 0.f), img_h);
                x1 = rintf(x1); y1 = rintf(y1); x2 = rintf(x2); y2 = rintf(y2);                                       // .round()
            }
            float* o = out + ((long long)b * max_det + k0 + q) * 6;
            o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = cd.conf; o[5] = cd.cls;
        }
        for (int w = blk + 1 + threadIdx.x; w < nw; w += blockDim.x) {  // suppress in the later blocks
            unsigned long long acc = 0;
            for (int q = 0; q < nk; ++q) acc |= mb[(long long)kept_rows[q] * words + w];
            remv[w] |= acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_keep = k0 + nk;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_count[b] = s_keep < max_det ? s_keep : max_det;
}

inline int grid_for(long long total, int block) { long long g = (total + block - 1) / block; return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g)); }
}  // namespace

extern "C" const char* b2t_detect_last_error(void) { return g_det_err.c_str(); }

extern "C" int b2t_image_reorg(const float* img, void* out, int B, int H, int W, void* stream) {
    if (!img || !out || (H & 1) || (W & 1)) return dfail(B2T_EINVAL, "b2t_image_reorg: bad arguments");
    const long long total = (long long)B * (H / 2) * (W / 2);
    image_reorg_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img, (__nv_bfloat16*)out, B, H, W);
    return dcheck("image_reorg");
}

extern "C" int b2t_upsample2x(const void* src, int src_pitch, int src_coff, void* dst, int dst_pitch, int dst_coff, int B, int H, int W,
                              int C, void* stream) {
    if (!src || !dst || C % 8 || src_pitch % 8 || dst_pitch % 8 || src_coff % 8 || dst_coff % 8) return dfail(B2T_EINVAL, "b2t_upsample2x: channels / pitches must be multiples of 8");
    const long long total = (long long)B * 4 * H * W * (C / 8);
    upsample2x_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, src_pitch, src_coff, (__nv_bfloat16*)dst,
                                                                                 dst_pitch, dst_coff, B, H, W, C);
    return dcheck("upsample2x");
}

extern "C" int b2t_spp_pool(void* buf, int pitch, int C, int B, int H, int W, void* stream) {
    if (!buf || C % 2 || pitch < 4 * C) return dfail(B2T_EINVAL, "b2t_spp_pool: bad arguments");
    const long long total = (long long)B * H * W * (C / 2);
    spp_pool_kernel<<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)buf, pitch, C, B, H, W);
    return dcheck("spp_pool");
}

extern "C" int b2t_detect_decode(const float* raw, int raw_pitch, float* pred, int B, int H, int W, int na, int no, long long level_off,
                                 long long n_total, float stride, const float* anchors_host, void* stream) {
    if (!raw || !pred || na != 3 || !anchors_host) return dfail(B2T_EINVAL, "b2t_detect_decode: bad arguments (na must be 3)");
    const long long total = (long long)B * H * W * na * no;
    detect_decode_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(raw, raw_pitch, pred, B, H, W, na, no, level_off, n_total, stride,
        anchors_host[0], anchors_host[1], anchors_host[2], anchors_host[3], anchors_host[4], anchors_host[5]);
    return dcheck("detect_decode");
}

extern "C" size_t b2t_nms_workspace_bytes(int B, int max_cand, int max_nms) {
    if (max_nms > max_cand) max_nms = max_cand;
    const size_t words = (size_t)(max_nms + 63) / 64;
    return (size_t)B * ((size_t)max_cand * (sizeof(Cand) + 4) + (size_t)max_nms * (sizeof(Cand) + 16) + (size_t)max_nms * words * 8) + (size_t)B * 4 + 4096;
}

extern "C" int b2t_nms(const float* pred, int B, int N, int no, float conf_thres, float iou_thres, int max_det, int max_nms, int max_cand,
                       int post, float gain, float padw, float padh, float img_w, float img_h, void* workspace, size_t workspace_bytes,
                       float* out, int* out_count, void* stream) {
    if (!pred || !workspace || !out || !out_count || B < 1 || N < 1 || no < 6 || max_det < 1 || max_nms < 1 || max_cand < 1)
        return dfail(B2T_EINVAL, "b2t_nms: bad arguments");
    if (max_nms > max_cand) max_nms = max_cand;
    if (workspace_bytes < b2t_nms_workspace_bytes(B, max_cand, max_nms)) return dfail(B2T_EINVAL, "b2t_nms: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const int words = (max_nms + 63) / 64;
    unsigned char* p = (unsigned char*)workspace;
    p = (unsigned char*)(((size_t)p + 255) / 256 * 256);
    int* count = (int*)p;                      p += ((size_t)B * 4 + 255) / 256 * 256;
    Cand* cand = (Cand*)p;                     p += ((size_t)B * max_cand * sizeof(Cand) + 255) / 256 * 256;
    int* rank = (int*)p;                       p += ((size_t)B * max_cand * 4 + 255) / 256 * 256;
    Cand* sorted = (Cand*)p;                   p += ((size_t)B * max_nms * sizeof(Cand) + 255) / 256 * 256;
    float4* sbox = (float4*)p;                 p += ((size_t)B * max_nms * 16 + 255) / 256 * 256;
    unsigned long long* mask = (unsigned long long*)p;
    cudaMemsetAsync(count, 0, (size_t)B * 4, s);
    nms_filter_kernel<<<dim3((N + 255) / 256, B), 256, 0, s>>>(pred, N, no, conf_thres, cand, count, max_cand);
    nms_rank_kernel<<<dim3((max_cand + 255) / 256, B), 256, 0, s>>>(cand, count, max_cand, rank);
    nms_scatter_kernel<<<dim3((max_cand + 255) / 256, B), 256, 0, s>>>(cand, count, rank, max_cand, max_nms, 4096.f, sbox, sorted);
    nms_mask_kernel<<<dim3(148 * 8, B), 64, 0, s>>>(sbox, count, max_cand, max_nms, iou_thres, mask, words);
    nms_select_kernel<<<B, 256, (size_t)words * 8, s>>>(sorted, mask, count, max_cand, max_nms, words, max_det, out, out_count, post, gain, padw,
                                                        padh, img_w, img_h);
    return dcheck("nms");
}
