// b2t_kalman.cuh -- constant-velocity Kalman filter, 8 lanes per track.
//
// Replaces tracker/kalman_filter.py (KalmanFilter :158-363, BoTSORTKalmanFilter :414-605,
// NSAKalmanFilter :607-646) and botsort.multi_gmc (tracker/botsort.py:250-269).
//
// Thread mapping: a warp carries 4 tracks; lane r = lane & 7 of a group owns mean[r] and row r
// of the 8x8 covariance (64 B contiguous in fp64 -> the warp's loads/stores are fully coalesced,
// 4 x 576 B).  Cross-row terms travel by width-8 shuffles.  Every lane of the warp must call
// these functions (inactive groups carry dummy data).
//
// T = double reproduces the reference's float64 arithmetic; the *_f32 flags reproduce the places
// where NumPy 2 keeps the reference's noise terms in float32 because the track's mean is still
// float32 (oracle/kalman.py).  T = float is the all-fp32 variant.
#pragma once
#include "b2t_prims.cuh"

namespace b2t {

enum { FMT_XYAH = 0, FMT_XYWH = 1, FMT_NSA = 2 };

template <class T> struct KRow {
    T m;      // mean[r]
    T p[8];   // cov[r][0..7]
};

template <class T> B2T_DEV void kf_load(KRow<T>& k, const T* mean, const T* cov, int r) {
    k.m = mean[r];
#pragma unroll
    for (int j = 0; j < 8; ++j) k.p[j] = cov[r * 8 + j];
}
template <class T> B2T_DEV void kf_store(const KRow<T>& k, T* mean, T* cov, int r) {
    mean[r] = k.m;
#pragma unroll
    for (int j = 0; j < 8; ++j) cov[r * 8 + j] = k.p[j];
}

template <class T> struct IsF32 { static const bool v = false; };
template <> struct IsF32<float> { static const bool v = true; };

// std-dev of one noise component; f32 = evaluate in float32 and return the float32 value.
template <class T> B2T_DEV T noise_std(T weight, T base, bool f32) {
    if (IsF32<T>::v || f32) return (T)((float)weight * (float)base);
    return weight * base;
}
// square; f32sq = the square itself is rounded to float32 (NumPy squares a float32 array).
template <class T> B2T_DEV T noise_var(T s, bool f32sq) {
    if (IsF32<T>::v || f32sq) { float f = (float)s; return (T)(f * f); }
    return s * s;
}

// Process noise Q[r][r].  kalman_filter.py:308-318 (xyah) / :550-560 (xywh).
template <class T> B2T_DEV T kf_q(int r, int fmt, T w, T h, bool f32) {
    const bool pos = r < 4;
    const int rr = r & 3;
    const T wgt = pos ? (T)(1.0 / 20) : (T)(1.0 / 160);
    T s;
    if (fmt == FMT_XYWH) {
        s = noise_std<T>(wgt, (rr & 1) ? h : w, f32);
    } else if (rr == 2) {
        s = pos ? (T)1e-2 : (T)1e-5;
        if (IsF32<T>::v || f32) s = (T)((float)s);
    } else {
        s = noise_std<T>(wgt, h, f32);
    }
    return noise_var<T>(s, f32);
}

// STrack.multi_predict + KalmanFilter.multi_predict (basetrack.py:253-271, kalman_filter.py:289-329).
//   zero_vh : state != Tracked -> mean[7] = 0 first (q6)
//   q_f32   : every mean of the batch is still float32 -> Q evaluated in float32
template <class T> B2T_DEV void kf_predict(KRow<T>& k, int r, int fmt, bool zero_vh, bool q_f32) {
    if (zero_vh && r == 7) k.m = (T)0;
    const T w = shfl(k.m, 2, 8), h = shfl(k.m, 3, 8);
    const int up = (r & 3) + 4;
    const T m_hi = shfl(k.m, up, 8);
    T lrow[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const T other = shfl(k.p[j], up, 8);
        lrow[j] = (r < 4) ? (k.p[j] + other) : k.p[j];      // (F P)[r][j]
    }
    if (r < 4) k.m = k.m + m_hi;                             // mean F^T
    const T q = kf_q<T>(r, fmt, w, h, q_f32);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        T v = (j < 4) ? (lrow[j] + lrow[(j + 4) & 7]) : lrow[j];   // (F P F^T)[r][j]
        if (j == r) v = v + q;
        k.p[j] = v;
    }
}

// Measurement noise R[c][c].  kalman_filter.py:277-282 / :522-527 / :617-626.
//   conf < 0 : no confidence argument (plain filters, and re_activate for NSA -- q7)
template <class T> B2T_DEV T kf_r(int c, int fmt, T w, T h, bool mean_f32, float conf) {
    const T wp = (T)(1.0 / 20);
    if (fmt == FMT_XYWH) {
        T s = noise_std<T>(wp, (c & 1) ? h : w, mean_f32);
        return noise_var<T>(s, mean_f32);
    }
    T s = (c == 2) ? (T)1e-1 : noise_std<T>(wp, h, mean_f32);
    if (fmt == FMT_NSA && conf >= 0.f) {
        const float omc = 1.0f - conf;                       // np.float32 scalar
        if (IsF32<T>::v) return noise_var<T>((T)(omc * (float)s), true);
        if (c == 2) {
            s = (T)(omc * 0.1f);                             // float32 * python float -> float32
            return noise_var<T>(s, mean_f32);               // all-float32 list when the mean is float32
        }
        if (mean_f32) return noise_var<T>((T)(omc * (float)s), true);
        s = (T)omc * s;
    }
    return noise_var<T>(s, false);
}

// KalmanFilter.project + update (kalman_filter.py:260-287, :331-363): Cholesky of the 4x4
// innovation covariance, gain by two triangular solves, mean / covariance correction in the
// association order NumPy uses (K (S K^T)).
template <class T> B2T_DEV void kf_update(KRow<T>& k, int r, int fmt, const T* z, bool mean_f32, float conf) {
    T S[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) S[a][b] = shfl(k.p[b], a, 8);
    const T w = shfl(k.m, 2, 8), h = shfl(k.m, 3, 8);
#pragma unroll
    for (int c = 0; c < 4; ++c) S[c][c] = S[c][c] + kf_r<T>(c, fmt, w, h, mean_f32, conf);
    // lower Cholesky (dpotrf order for n = 4); one reciprocal per pivot instead of 14 divisions
    // (an fp64 division is ~60 SASS instructions) -- within an ulp or two of the LAPACK result.
    T L[4][4], inv[4];
    L[0][0] = sqrt(S[0][0]);                                   inv[0] = (T)1 / L[0][0];
    L[1][0] = S[1][0] * inv[0];
    L[2][0] = S[2][0] * inv[0];
    L[3][0] = S[3][0] * inv[0];
    L[1][1] = sqrt(S[1][1] - L[1][0] * L[1][0]);               inv[1] = (T)1 / L[1][1];
    L[2][1] = (S[2][1] - L[2][0] * L[1][0]) * inv[1];
    L[3][1] = (S[3][1] - L[3][0] * L[1][0]) * inv[1];
    L[2][2] = sqrt((S[2][2] - L[2][0] * L[2][0]) - L[2][1] * L[2][1]);   inv[2] = (T)1 / L[2][2];
    L[3][2] = ((S[3][2] - L[3][0] * L[2][0]) - L[3][1] * L[2][1]) * inv[2];
    L[3][3] = sqrt(((S[3][3] - L[3][0] * L[3][0]) - L[3][1] * L[3][1]) - L[3][2] * L[3][2]);   inv[3] = (T)1 / L[3][3];
    // gain row r: solve S g = P[r][0:4]^T  (forward, then backward substitution)
    T g[4];
    g[0] = k.p[0] * inv[0];
    g[1] = (k.p[1] - L[1][0] * g[0]) * inv[1];
    g[2] = ((k.p[2] - L[2][0] * g[0]) - L[2][1] * g[1]) * inv[2];
    g[3] = (((k.p[3] - L[3][0] * g[0]) - L[3][1] * g[1]) - L[3][2] * g[2]) * inv[3];
    g[3] = g[3] * inv[3];
    g[2] = (g[2] - L[3][2] * g[3]) * inv[2];
    g[1] = ((g[1] - L[2][1] * g[2]) - L[3][1] * g[3]) * inv[1];
    g[0] = (((g[0] - L[1][0] * g[1]) - L[2][0] * g[2]) - L[3][0] * g[3]) * inv[0];
    // innovation and mean
    T acc = (T)0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const T inn = z[c] - shfl(k.m, c, 8);
        acc = acc + inn * g[c];
    }
    k.m = k.m + acc;
    // (S K^T)[a][r]
    T skt[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) skt[a] = ((S[a][0] * g[0] + S[a][1] * g[1]) + S[a][2] * g[2]) + S[a][3] * g[3];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        T d = (T)0;
#pragma unroll
        for (int a = 0; a < 4; ++a) d = d + g[a] * shfl(skt[a], j, 8);
        k.p[j] = k.p[j] - d;
    }
}

// KalmanFilter.initiate (kalman_filter.py:190-221 / :435-466); z holds float32 values.
template <class T> B2T_DEV void kf_initiate(KRow<T>& k, int r, int fmt, const T* z) {
    k.m = (r < 4) ? z[r] : (T)0;
    const bool pos = r < 4;
    const int rr = r & 3;
    T var;
    if (fmt == FMT_XYWH) {
        const float s = (pos ? 0.1f : 0.0625f) * (float)((rr & 1) ? z[3] : z[2]);
        var = (T)(s * s);                                   // float32 list -> float32 squares
    } else if (rr == 2) {
        const T c = pos ? (T)1e-2 : (T)1e-5;
        var = IsF32<T>::v ? (T)((float)c * (float)c) : c * c;
    } else {
        const float s = (pos ? 0.1f : 0.0625f) * (float)z[3];
        var = IsF32<T>::v ? (T)(s * s) : (T)s * (T)s;       // float32 std squared in float64
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) k.p[j] = (j == r) ? var : (T)0;
}

// multi_gmc (botsort.py:250-269): mean <- R8 mean (+t), P <- R8 P R8^T, R8 = kron(I4, A), A = H[:, :2].
// warp6 = {a00, a01, tx, a10, a11, ty}.
template <class T> B2T_DEV void kf_gmc(KRow<T>& k, int r, const T* warp6) {
    const int odd = r & 1;
    const T ra = odd ? warp6[3] : warp6[0];   // A[r%2][0]
    const T rb = odd ? warp6[4] : warp6[1];   // A[r%2][1]
    const T m_other = shfl_xor(k.m, 1, 8);
    const T m_even = odd ? m_other : k.m, m_odd = odd ? k.m : m_other;
    T nm = ra * m_even + rb * m_odd;
    if (r == 0) nm = nm + warp6[2];
    if (r == 1) nm = nm + warp6[5];
    T mrow[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const T other = shfl_xor(k.p[j], 1, 8);
        const T pe = odd ? other : k.p[j], po = odd ? k.p[j] : other;
        mrow[j] = ra * pe + rb * po;                         // (R8 P)[r][j]
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const T ca = (j & 1) ? warp6[3] : warp6[0];
        const T cb = (j & 1) ? warp6[4] : warp6[1];
        k.p[j] = mrow[j & 6] * ca + mrow[(j & 6) | 1] * cb;  // (. R8^T)[r][j]
    }
    k.m = nm;
}

// ---- box conversions -------------------------------------------------------------------
// Detection tlbr (float32) -> measurement, float32 arithmetic as in basetrack.py:111-150.
template <class T> B2T_DEV void det_to_meas(int fmt, float x1, float y1, float x2, float y2, T* z) {
    const float w = x2 - x1, h = y2 - y1;
    if (fmt == FMT_XYWH) {
        z[0] = (T)(x1 + floorf(w / 2.0f));
        z[1] = (T)(y1 + floorf(h / 2.0f));
        z[2] = (T)w;
        z[3] = (T)h;
    } else {
        z[0] = (T)(x1 + w / 2.0f);
        z[1] = (T)(y1 + h / 2.0f);
        z[2] = (T)(w / h);
        z[3] = (T)h;
    }
}

// STrack.tlwh / .tlbr from the state mean (basetrack.py:183-219), in the mean's dtype.
template <class T> B2T_DEV void mean_to_tlwh(int fmt, const T* m, bool mean_f32, T* o) {
    if (mean_f32 && !IsF32<T>::v) {
        float w = (float)m[2], h = (float)m[3];
        if (fmt != FMT_XYWH) w = w * h;
        o[0] = (T)((float)m[0] - w / 2.0f);
        o[1] = (T)((float)m[1] - h / 2.0f);
        o[2] = (T)w;
        o[3] = (T)h;
    } else {
        T w = m[2], h = m[3];
        if (fmt != FMT_XYWH) w = w * h;
        o[0] = m[0] - w / (T)2;
        o[1] = m[1] - h / (T)2;
        o[2] = w;
        o[3] = h;
    }
}
template <class T> B2T_DEV void mean_to_tlbr(int fmt, const T* m, bool mean_f32, T* o) {
    mean_to_tlwh<T>(fmt, m, mean_f32, o);
    if (mean_f32 && !IsF32<T>::v) {
        o[2] = (T)((float)o[2] + (float)o[0]);
        o[3] = (T)((float)o[3] + (float)o[1]);
    } else {
        o[2] = o[2] + o[0];
        o[3] = o[3] + o[1];
    }
}

}  // namespace b2t
