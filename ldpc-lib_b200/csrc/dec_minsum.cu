// Table-driven parity kernels of the min-sum family: LMS_DEC, MS_DEC (double or float) and IMS_DEC.
// One frame per CTA at a time, persistent grid, state in the CTA's global workspace slice.  These
// kernels reproduce the reference's arithmetic order exactly (compiled with -fmad=false) and run for
// every code; the shared-memory throughput kernels live in lms_fast.cu / ims_fast.cu.
//
//   LMS_DEC  lmin_sum_decod_qc_lm   decoders.cpp:5064-5425   layered offset min-sum, beta = 0.4
//   MS_DEC   min_sum_decod_qc_lm    decoders.cpp:4554-4767   flooding normalised min-sum
//   IMS_DEC  imin_sum_decod_qc_lm   decoders.cpp:5430-5690   flooding fixed-point min-sum
//
// Lane formulation: row n of block row j touches bit col*Z + (n + shift) mod Z; lanes of one block
// row touch disjoint bits, so they run in parallel.  Orders that matter for rounding / saturation are
// kept: edges of a row by ascending block column, edges of a column by ascending block row.
#include "dec_common.cuh"

namespace ldpcb200 {

template <typename T>
__device__ __forceinline__ int syndrome_neg(const QcDev& g, const T* soft)
{
    int bad = 0;
    for (int r = threadIdx.x; r < g.R; r += blockDim.x) {
        int j = r / g.Z, n = r - j * g.Z, s = 0;
        for (int e = g.rp[j]; e < g.rp[j + 1]; e++)
            s ^= soft[g.col[e] * g.Z + wrapz(n + g.sh[e], g.Z)] < 0;
        bad |= s;
    }
    return __syncthreads_or(bad);
}

// ------------------------------------------------------------------------------------------------
template <typename T>
struct LmsGeneric {
    static size_t ws_bytes(const QcHost& g, int nt)
    {
        return carve_bytes(g.N, sizeof(T)) + 2 * carve_bytes(g.R, sizeof(T)) + carve_bytes(g.R, 4) +
               carve_bytes((size_t)g.E * g.Z, 1) + carve_bytes((size_t)g.maxdeg * nt, sizeof(T));
    }
    static __device__ void frame(const QcDev& g, const DecParams&, const FrameIO& io, int f, char* ws)
    {
        const int Z = g.Z, N = g.N, R = g.R, nt = blockDim.x, tid = threadIdx.x;
        T* soft = carve<T>(ws, N);
        T* min1 = carve<T>(ws, R);
        T* min2 = carve<T>(ws, R);
        int* ps = carve<int>(ws, R);                    // pos | sign << 16
        uint8_t* esign = carve<uint8_t>(ws, (size_t)g.E * Z);
        T* vbuf = carve<T>(ws, (size_t)g.maxdeg * nt);
        const T beta = (T)0.4;                          // decoders.cpp:5163 (the beta argument is ignored)
        const T MAXV = (T)32767;                        // MAX_VAL, decoders.cpp:4301

        for (int i = tid; i < N; i += nt) soft[i] = (T)load_llr(io, N, f, i);
        for (int i = tid; i < R; i += nt) { min1[i] = 0; min2[i] = 0; ps[i] = 0; }
        for (int i = tid; i < g.E * Z; i += nt) esign[i] = 0;
        __syncthreads();

        int parity = syndrome_neg(g, soft);             // :5111-5115
        int ret = 0, locked = 0, iter;
        const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
        if (!parity) { ret = 1; locked = 1; }           // already a codeword: 0 + 1
        for (iter = 0; iter < io.maxiter; iter++) {
            if (!parity && !noexit) break;              // :5119
            for (int j = 0; j < g.b; j++) {
                const int e0 = g.rp[j], deg = g.rp[j + 1] - e0;
                for (int n = tid; n < Z; n += nt) {
                    const int r = j * Z + n;
                    const T pm1 = min1[r], pm2 = min2[r];
                    const int ppos = ps[r] & 0xffff, psign = ps[r] >> 16;
                    T c1 = MAXV, c2 = MAXV;
                    int cpos = 0, csign = 0;
                    for (int q = 0; q < deg; q++) {
                        const int e = e0 + q, k = g.col[e];
                        const int idx = k * Z + wrapz(n + g.sh[e], Z);
                        T pabs = ppos == k ? pm2 : pm1;                          // :5152
                        int psgn = esign[(size_t)e * Z + n] ^ psign;             // :5156
                        T pval = psgn ? -pabs : pabs;
                        T v = soft[idx] - pval;                                  // :5158
                        int s = v < 0;
                        T a = v < (T)0 ? -v : v;
                        a -= beta;                                               // :5166
                        a = a < 0 ? (T)0 : a;                                    // :5168
                        vbuf[(size_t)q * nt + tid] = v;
                        esign[(size_t)e * Z + n] = (uint8_t)s;
                        csign ^= s;
                        if (a < c1) { cpos = k; c2 = c1; c1 = a; }               // process_check_node :5012-5027
                        else if (a < c2) c2 = a;
                    }
                    min1[r] = c1; min2[r] = c2; ps[r] = cpos | (csign << 16);
                    for (int q = 0; q < deg; q++) {
                        const int e = e0 + q, k = g.col[e];
                        const int idx = k * Z + wrapz(n + g.sh[e], Z);
                        T cabs = cpos == k ? c2 : c1;                            // :5193
                        T cval = (esign[(size_t)e * Z + n] ^ csign) ? -cabs : cabs;
                        soft[idx] = vbuf[(size_t)q * nt + tid] + cval;           // :5199-5204
                    }
                }
                __syncthreads();
            }
            parity = syndrome_neg(g, soft);             // :5281-5284
            if (!parity && !locked) { ret = iter + 1; locked = 1; }
            if (!parity && !noexit) break;
        }
        if (!locked) ret = parity ? -iter : iter + 1;   // :5424
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, soft[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(soft[i] < 0); });   // :5421
    }
};

// ------------------------------------------------------------------------------------------------
template <typename T>
struct MsGeneric {
    static size_t ws_bytes(const QcHost& g, int)
    {
        return 2 * carve_bytes(g.N, sizeof(T)) + 2 * carve_bytes(g.R, sizeof(T)) + carve_bytes(g.R, 4) +
               carve_bytes((size_t)g.E * g.Z, 1);
    }
    static __device__ void frame(const QcDev& g, const DecParams& dp, const FrameIO& io, int f, char* ws)
    {
        const int Z = g.Z, N = g.N, R = g.R, nt = blockDim.x, tid = threadIdx.x;
        T* soft = carve<T>(ws, N);
        T* y = carve<T>(ws, N);
        T* min1 = carve<T>(ws, R);
        T* min2 = carve<T>(ws, R);
        int* ps = carve<int>(ws, R);
        uint8_t* esign = carve<uint8_t>(ws, (size_t)g.E * Z);
        const T MAXV = (T)32767;
        const T alpha = (T)dp.alpha;

        for (int i = tid; i < N; i += nt) { y[i] = (T)load_llr(io, N, f, i); soft[i] = y[i]; }
        for (int i = tid; i < R; i += nt) { min1[i] = 0; min2[i] = 0; ps[i] = 0; }   // :4579-4585
        for (int i = tid; i < g.E * Z; i += nt) esign[i] = 0;                        // :4596
        __syncthreads();

        int parity = 1, ret = 0, locked = 0, iter;
        const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
        for (iter = 0; iter < io.maxiter; iter++) {
            // STATE 1 + 2 (:4633-4685): per variable, sum over the column's block rows ascending
            for (int v = tid; v < N; v += nt) {
                const int i = v / Z, k = v - i * Z;
                T acc = 0;
                for (int q = g.cp[i]; q < g.cp[i + 1]; q++) {
                    const int e = g.cedge[q], j = g.row[e];
                    const int n = wrapz(k - g.sh[e] + Z, Z);
                    const int r = j * Z + n;
                    T tmp = (ps[r] & 0xffff) == i ? min2[r] : min1[r];           // :4649
                    T val = (esign[(size_t)e * Z + n] ^ (ps[r] >> 16)) ? -tmp : tmp;
                    acc = acc + val;                                             // :4658
                }
                soft[v] = y[v] + acc * alpha;                                    // :4682
            }
            __syncthreads();
            // STATE 3 (:4688-4755)
            int bad = 0;
            for (int r = tid; r < R; r += nt) {
                const int j = r / Z, n = r - j * Z;
                const T pm1 = min1[r], pm2 = min2[r];
                const int ppos = ps[r] & 0xffff, psign = ps[r] >> 16;
                T c1 = MAXV, c2 = MAXV;
                int cpos = 0, csign = 0, synd = 0;
                for (int e = g.rp[j]; e < g.rp[j + 1]; e++) {
                    const int k = g.col[e];
                    T rs = soft[k * Z + wrapz(n + g.sh[e], Z)];
                    synd ^= rs < 0;                                              // :4711
                    T old = ppos == k ? pm2 : pm1;                               // :4714
                    T val = old * alpha;                                         // :4719
                    T tt = (esign[(size_t)e * Z + n] ^ psign) ? -val : val;
                    tt = rs - tt;                                                // :4722
                    int s = tt < 0;
                    esign[(size_t)e * Z + n] = (uint8_t)s;
                    csign ^= s;
                    val = tt < (T)0 ? -tt : tt;                                  // :4729
                    val = val > MAXV ? MAXV : val;                               // :4730
                    if (val < c1) { cpos = k; c2 = c1; c1 = val; }
                    else if (val < c2) c2 = val;
                }
                min1[r] = c1; min2[r] = c2; ps[r] = cpos | (csign << 16);
                bad |= synd;
            }
            parity = __syncthreads_or(bad);
            if (!parity) { if (!locked) { ret = iter + 1; locked = 1; } if (!noexit) break; }   // :4761
        }
        if (!locked) ret = parity ? -iter : iter + 1;                            // :4766
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, soft[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(soft[i] < 0); });
    }
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sat_s16(int x, int mx) { return x > mx ? mx : (x < -mx ? -mx : x); }   // limit_val :4308

struct ImsGeneric {
    static size_t ws_bytes(const QcHost& g, int)
    {
        return 2 * carve_bytes(g.N, 2) + 2 * carve_bytes(g.R, 2) + carve_bytes(g.R, 4) +
               carve_bytes((size_t)g.E * g.Z, 1);
    }
    static __device__ void frame(const QcDev& g, const DecParams& dp, const FrameIO& io, int f, char* ws)
    {
        const int Z = g.Z, N = g.N, R = g.R, nt = blockDim.x, tid = threadIdx.x;
        int16_t* soft = carve<int16_t>(ws, N);
        int16_t* iy = carve<int16_t>(ws, N);
        int16_t* min1 = carve<int16_t>(ws, R);
        int16_t* min2 = carve<int16_t>(ws, R);
        int* ps = carve<int>(ws, R);
        uint8_t* esign = carve<uint8_t>(ws, (size_t)g.E * Z);
        const int max_data = (int16_t)((1L << (dp.dbits - 1)) - 1);              // :5445
        const int max_quant = (int16_t)((1L << (dp.qbits - 1)) - 1);             // :5446
        const int ialpha = (int)(dp.alpha * (1L << 4));                          // MS_ALPHA_FPP = 4, :5458
        const double thr = dp.thr;

        // per-frame energy: en += y[i]*y[i], i ascending, in double (:5476-5477).  The products are
        // formed in parallel, the additions run in the reference's order on one thread.
        __shared__ double s_sq[512];
        __shared__ double s_coef;
        if (io.coef) {
            if (tid == 0) s_coef = io.coef[f];                                   // from the energy pre-pass (channel.cu)
        } else {
            double en = 0;
            for (int base = 0; base < N; base += 512) {
                __syncthreads();
                for (int i = tid; i < 512 && base + i < N; i += nt) {
                    double v = load_llr(io, N, f, base + i);
                    s_sq[i] = v * v;
                }
                __syncthreads();
                if (tid == 0) {
                    int m = min(512, N - base);
                    for (int i = 0; i < m; i++) en += s_sq[i];
                }
            }
            if (tid == 0) s_coef = sqrt(N / en);                                 // :5479
        }
        __syncthreads();
        const double coef = s_coef;
        for (int i = tid; i < N; i += nt) {                                      // :5481-5499
            double val = load_llr(io, N, f, i);
            int sign = 0;
            if (val < 0) { val = -val; sign = 1; }
            val *= coef;
            if (val > thr) val = thr;
            int ival = (int16_t)floor(val * max_quant / thr + 0.5);
            iy[i] = (int16_t)(sign ? -ival : ival);
            soft[i] = iy[i];                                            // defined results when no iteration runs (maxiter == 0)
            if (io.aux) io.aux[(size_t)f * N + i] = iy[i];
        }
        for (int i = tid; i < R; i += nt) { min1[i] = 0; min2[i] = 0; ps[i] = 0; }
        for (int i = tid; i < g.E * Z; i += nt) esign[i] = 0;
        __syncthreads();

        int parity = 1, ret = 0, locked = 0, iter;
        const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
        for (iter = 0; iter < io.maxiter; iter++) {
            // STATE 1 + 2 (:5536-5603): saturate after EVERY add, block rows ascending
            for (int v = tid; v < N; v += nt) {
                const int i = v / Z, k = v - i * Z;
                int acc = 0;
                for (int q = g.cp[i]; q < g.cp[i + 1]; q++) {
                    const int e = g.cedge[q], j = g.row[e];
                    const int n = wrapz(k - g.sh[e] + Z, Z);
                    const int r = j * Z + n;
                    int tmp = (ps[r] & 0xffff) == i ? min2[r] : min1[r];
                    tmp = (int16_t)((tmp * ialpha) >> 4);                        // :5554
                    int val = (esign[(size_t)e * Z + n] ^ (ps[r] >> 16)) ? -tmp : tmp;
                    acc = sat_s16((int16_t)(acc + val), max_data);               // :5567-5568
                }
                soft[v] = (int16_t)sat_s16((int16_t)(iy[v] + acc), max_data);    // :5599-5601
            }
            __syncthreads();
            // STATE 3 (:5608-5678)
            int bad = 0;
            for (int r = tid; r < R; r += nt) {
                const int j = r / Z, n = r - j * Z;
                const int pm1 = min1[r], pm2 = min2[r];
                const int ppos = ps[r] & 0xffff, psign = ps[r] >> 16;
                int c1 = max_data, c2 = max_data, cpos = 0, csign = 0, synd = 0;
                for (int e = g.rp[j]; e < g.rp[j + 1]; e++) {
                    const int k = g.col[e];
                    int rs = soft[k * Z + wrapz(n + g.sh[e], Z)];
                    synd ^= rs < 0;                                              // :5631
                    int old = ppos == k ? pm2 : pm1;
                    int val = (int16_t)((old * ialpha) >> 4);                    // :5640
                    int tt = (esign[(size_t)e * Z + n] ^ psign) ? -val : val;
                    int v2c = (int16_t)(rs - tt);                                // :5646
                    int s = v2c < 0;
                    esign[(size_t)e * Z + n] = (uint8_t)s;
                    csign ^= s;
                    val = (int16_t)(v2c < 0 ? -v2c : v2c);
                    val = val > max_data ? max_data : val;                       // :5653
                    if (val < c1) { cpos = k; c2 = c1; c1 = val; }
                    else if (val < c2) c2 = val;
                }
                min1[r] = (int16_t)c1; min2[r] = (int16_t)c2; ps[r] = cpos | (csign << 16);
                bad |= synd;
            }
            parity = __syncthreads_or(bad);
            if (!parity) { if (!locked) { ret = iter + 1; locked = 1; } if (!noexit) break; }
        }
        if (!locked) ret = parity ? -iter : iter + 1;                            // :5689
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, soft[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(soft[i] < 0); });
    }
};

// ------------------------------------------------------------------------------------------------
template <class Dec>
__global__ void __launch_bounds__(512) generic_minsum_kernel(QcDev g, DecParams dp, FrameIO io, char* ws, size_t ws_stride)
{
    // the decoder state lives in shared memory when it fits (ws == nullptr), else in this CTA's slice of an
    // L2-resident global workspace
    extern __shared__ __align__(16) char dyn_ws[];
    char* w = ws ? ws + (size_t)blockIdx.x * ws_stride : dyn_ws;
    for (;;) {
        int f = next_frame(io);
        if (f >= io.nf) break;
        Dec::frame(g, dp, io, f, w);
    }
}

template <class Kern>
static void launch_one(Kern kern, const QcDev& g, const DecParams& dp, const FrameIO& io, char* ws, size_t ws_stride, size_t smem_ws,
                       int grid, int nt, cudaStream_t s)
{
    if (smem_ws) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws);
    kern<<<grid, nt, smem_ws, s>>>(g, dp, io, smem_ws ? nullptr : ws, ws_stride);
}

size_t minsum_workspace_bytes(int decoder_id, int precision, const QcHost& g, int nt)
{
    switch (decoder_id) {
    case LDPCB200_LMS_DEC: return precision == 32 ? LmsGeneric<float>::ws_bytes(g, nt) : LmsGeneric<double>::ws_bytes(g, nt);
    case LDPCB200_MS_DEC:  return precision == 32 ? MsGeneric<float>::ws_bytes(g, nt) : MsGeneric<double>::ws_bytes(g, nt);
    case LDPCB200_IMS_DEC: return ImsGeneric::ws_bytes(g, nt);
    }
    return 0;
}

cudaError_t launch_minsum_generic(int decoder_id, int precision, const QcDev& g, const DecParams& dp,
                                  const FrameIO& io, char* ws, size_t ws_stride, size_t smem_ws, int grid, int nt, cudaStream_t s)
{
    switch (decoder_id) {
    case LDPCB200_LMS_DEC:
        if (precision == 32) launch_one(generic_minsum_kernel<LmsGeneric<float>>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s);
        else launch_one(generic_minsum_kernel<LmsGeneric<double>>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s);
        break;
    case LDPCB200_MS_DEC:
        if (precision == 32) launch_one(generic_minsum_kernel<MsGeneric<float>>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s);
        else launch_one(generic_minsum_kernel<MsGeneric<double>>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s);
        break;
    case LDPCB200_IMS_DEC:
        launch_one(generic_minsum_kernel<ImsGeneric>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s);
        break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace ldpcb200
