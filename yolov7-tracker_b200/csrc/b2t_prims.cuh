// b2t_prims.cuh -- warp / block building blocks shared by the tracker kernels.
#pragma once
#include "b2t_platform.cuh"

namespace b2t {

B2T_DEV int lane_id() { return (int)(threadIdx.x & 31u); }
B2T_DEV int warp_id() { return (int)(threadIdx.x >> 5); }
B2T_DEV int num_warps() { return (int)(blockDim.x >> 5); }
B2T_DEV unsigned lanemask_lt() { return (1u << lane_id()) - 1u; }

template <class T> B2T_DEV T shfl(T v, int src, int width = 32) { return __shfl_sync(B2T_FULL, v, src, width); }
template <class T> B2T_DEV T shfl_up(T v, int d) { return __shfl_up_sync(B2T_FULL, v, (unsigned)d); }
template <class T> B2T_DEV T shfl_xor(T v, int m, int width = 32) { return __shfl_xor_sync(B2T_FULL, v, m, width); }

template <class T> struct Inf;
template <> struct Inf<float> { B2T_DEV static float v() { return __int_as_float(0x7f800000); } };
template <> struct Inf<double> { B2T_DEV static double v() { return 1.0e300 * 1.0e300; } };

// Bump allocator over the dynamic shared memory block.  Every thread performs the same
// arithmetic, so no synchronisation is involved.
struct Arena {
    unsigned char* base;
    size_t off;
    B2T_DEV Arena(unsigned char* b) : base(b), off(0) {}
    template <class T> B2T_DEV T* take(int n) {
        off = (off + 15) & ~size_t(15);
        T* p = reinterpret_cast<T*>(base + off);
        off += sizeof(T) * (size_t)(n > 0 ? n : 1);
        return p;
    }
};
// Host-side mirror of Arena::take for sizing the launch.
struct ArenaSize {
    size_t off = 0;
    template <class T> void take(int n) { off = (off + 15) & ~size_t(15); off += sizeof(T) * (size_t)(n > 0 ? n : 1); }
};

// In-place exclusive scan of a[0..n) (shared memory), all threads of the CTA participate.
// scratch: >= 33 ints of shared memory.  Returns the total.
B2T_DEV int block_exscan(int* a, int n, int* scratch) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int per = (n + nt - 1) / nt;
    int s = tid * per, e = s + per;
    if (s > n) s = n;
    if (e > n) e = n;
    int sum = 0;
    for (int i = s; i < e; ++i) sum += a[i];
    int inc = sum;
    for (int d = 1; d < 32; d <<= 1) {
        int t = shfl_up(inc, d);
        if (lane_id() >= d) inc += t;
    }
    if (lane_id() == 31) scratch[warp_id()] = inc;
    __syncthreads();
    if (warp_id() == 0) {
        int w = lane_id() < num_warps() ? scratch[lane_id()] : 0;
        int winc = w;
        for (int d = 1; d < 32; d <<= 1) {
            int t = shfl_up(winc, d);
            if (lane_id() >= d) winc += t;
        }
        scratch[lane_id()] = winc - w;
        if (lane_id() == 31) scratch[32] = winc;
    }
    __syncthreads();
    int base = scratch[warp_id()] + inc - sum;
    for (int i = s; i < e; ++i) { int t = a[i]; a[i] = base; base += t; }
    int total = scratch[32];
    __syncthreads();
    return total;
}

// Order-preserving compaction: out[] receives every i in [0, n) with pred(i), ascending.
// All threads participate; scratch: >= 32 ints.  Returns the count (uniform).
template <class Pred> B2T_DEV int block_compact(int n, Pred pred, int* out, int* scratch) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    int base = 0;
    for (int c0 = 0; c0 < n; c0 += nt) {
        const int i = c0 + tid;
        const bool p = (i < n) && pred(i);
        const unsigned bal = __ballot_sync(B2T_FULL, p);
        if (lane_id() == 0) scratch[warp_id()] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < num_warps(); ++w) {
            int c = scratch[w];
            if (w < warp_id()) woff += c;
            tot += c;
        }
        if (p) out[base + woff + __popc(bal & lanemask_lt())] = i;
        base += tot;
        __syncthreads();
    }
    return base;
}

}  // namespace b2t
