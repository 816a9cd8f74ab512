// a8: the L-BFGS driver of the PLM fit, resident on the device (SURVEY.md 8a row a8, 8b `evc_plm_fit`).
//
// plmc minimises the objective with libLBFGS on the host CPU (reference call site
// evcouplings/couplings/tools.py:226-228 passes the iteration cap `-m`).  Here every n-vector (x, g, search
// direction, m correction pairs) lives in HBM inside the handle; per objective evaluation the host sees six
// doubles (one 48-byte D2H + one stream synchronisation) -- the scalars the More-Thuente line search decides on.
//
//   trial point      x_try = x + t d                                    (1 kernel, 3 vector passes)
//   objective        evc_plm_eval_data (expand -> GEMM -> softmax -> GEMM -> symmetrise)
//   multi-rank       -loglk is packed as three exact fixed-point limbs behind the gradient so that ONE
//                    all-reduce (callback; NCCL in the Python host) carries [g, fx]; every partial sum of a
//                    limb is an integer < 2^24, i.e. the fp32 reduction is exact and order-independent
//   regulariser      g += 2 lambda x fused with the five reductions lambda|x|^2, g.d, g.g, |h|^2, |J|^2
//   two-loop         2*bound+1 fused kernels "d += c v; partial(u.d)" (4 vector passes each) with the
//                    coefficients alpha/beta kept on the device
// All reductions use a fixed grid and a fixed tree => every rank of a data-parallel run takes bit-identical
// decisions without broadcasting anything.
//
// The line search is the safeguarded cubic/quadratic interpolation of More & Thuente (1994) with libLBFGS's
// default constants (ftol 1e-4, gtol 0.9, xtol 1e-7, 40 trials), first step 1/|g|, then 1.
#include <chrono>
#include <cmath>

#include "../../include/evcplm.h"
#include "common.cuh"
#include "internal.h"

namespace evc {

constexpr int FIT_BLOCKS = 1184;     // 8 CTAs of 256 threads per SM on 148 SMs
constexpr int FIT_THREADS = 256;
constexpr int FIT_NRED = 5;          // reductions of the regulariser kernel
constexpr int64_t FX_LIMB_BITS = 18;
constexpr double FX_SCALE = 65536.0; // fixed-point resolution 2^-16 of the packed -loglk

// device scalar block (doubles)
enum { SC_FX = 0, SC_NLL, SC_DG, SC_GG, SC_XXH, SC_XXJ, SC_YY, SC_COEF, SC_YS = 8 /* [m] */, SC_ALPHA = 8 + 32 /* [m] */, SC_COUNT = 8 + 64 };

__device__ __forceinline__ double fit_block_sum(double v, double *s_red)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += s_red[w];
    return tot;   // valid on thread 0
}

// fixed-tree sum of FIT_BLOCKS partials by one CTA of 1024 threads; result valid on thread 0
__device__ __forceinline__ double fit_final_sum(const double *__restrict__ partial, double *s_red)
{
    const int tid = threadIdx.x;
    double v = 0.0;
    for (int e = tid; e < FIT_BLOCKS; e += 1024) v += partial[e];
    __syncthreads();
    s_red[tid] = v;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) s_red[tid] += s_red[tid + o];
        __syncthreads();
    }
    return s_red[0];
}

__global__ void fit_step_kernel(float *__restrict__ xt, const float *__restrict__ x, const float *__restrict__ d,
                                float t, int64_t n)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        xt[e] = fmaf(t, d[e], x[e]);
}

// -loglk -> three fixed-point limbs (floats holding integers < 2^18) behind the gradient
__global__ void fit_pack_fx_kernel(const double *__restrict__ fx, float *__restrict__ limbs)
{
    double v = fx[0] * FX_SCALE;
    const double lim = 9.0e15;                      // |q| < 2^53: the top limb stays below 2^17 per rank (exact sums up to 64 ranks)
    v = fmin(fmax(v, -lim), lim);
    const long long q = llrint(v);
    const long long mask = (1ll << FX_LIMB_BITS) - 1;
    limbs[0] = (float)(q & mask);
    limbs[1] = (float)((q >> FX_LIMB_BITS) & mask);
    limbs[2] = (float)(q >> (2 * FX_LIMB_BITS));    // arithmetic shift keeps the sign
    limbs[3] = 0.f;
}

// g += 2 lambda x; partials of {lambda |x|^2, g.d, g.g, |h|^2, |J|^2}   (d may be null)
__global__ void fit_reg_dots_kernel(const float *__restrict__ x, float *__restrict__ g, const float *__restrict__ d,
                                    int64_t n, int64_t nh, float lambda_h, float lambda_J,
                                    double *__restrict__ partial)
{
    __shared__ double s_red[32];
    double a_reg = 0.0, a_dg = 0.0, a_gg = 0.0, a_h = 0.0, a_J = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const bool is_h = e < nh;
        const float lam = is_h ? lambda_h : lambda_J;
        const float xv = x[e];
        const float gv = g[e] + 2.f * lam * xv;
        g[e] = gv;
        const double xx = (double)xv * (double)xv;
        a_reg += (double)lam * xx;
        if (is_h) a_h += xx; else a_J += xx;
        a_gg += (double)gv * (double)gv;
        if (d != nullptr) a_dg += (double)gv * (double)d[e];
    }
    double r[FIT_NRED] = {a_reg, a_dg, a_gg, a_h, a_J};
#pragma unroll
    for (int k = 0; k < FIT_NRED; k++) {
        const double tot = fit_block_sum(r[k], s_red);
        if (threadIdx.x == 0) partial[k * FIT_BLOCKS + blockIdx.x] = tot;
    }
}

__global__ void __launch_bounds__(1024)
fit_reg_final_kernel(const double *__restrict__ partial, const double *__restrict__ fx_data,
                     const float *__restrict__ limbs, double *__restrict__ sc)
{
    __shared__ double s_red[1024];
    double out[FIT_NRED];
    for (int k = 0; k < FIT_NRED; k++) out[k] = fit_final_sum(partial + k * FIT_BLOCKS, s_red);
    if (threadIdx.x == 0) {
        double nll;
        if (limbs != nullptr) {
            const long long q = (long long)limbs[0] + ((long long)limbs[1] << FX_LIMB_BITS) +
                                ((long long)limbs[2]) * (1ll << (2 * FX_LIMB_BITS));
            nll = (double)q / FX_SCALE;
        } else {
            nll = fx_data[0];
        }
        sc[SC_NLL] = nll;
        sc[SC_FX] = nll + out[0];
        sc[SC_DG] = out[1];
        sc[SC_GG] = out[2];
        sc[SC_XXH] = out[3];
        sc[SC_XXJ] = out[4];
    }
}

// plain dot (used for g.d at the start of a line search)
__global__ void fit_dot_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t n,
                               double *__restrict__ partial)
{
    __shared__ double s_red[32];
    double acc = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        acc += (double)a[e] * (double)b[e];
    const double tot = fit_block_sum(acc, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// mode 0: out = sum; 1: out = sum / den[0]; 2: out = aux[0] - sum / den[0]
__global__ void __launch_bounds__(1024)
fit_scalar_final_kernel(const double *__restrict__ partial, int mode, const double *__restrict__ den,
                        const double *__restrict__ aux, double *__restrict__ out)
{
    __shared__ double s_red[1024];
    const double v = fit_final_sum(partial, s_red);
    if (threadIdx.x == 0) {
        if (mode == 0) out[0] = v;
        else if (mode == 1) out[0] = v / den[0];
        else out[0] = aux[0] - v / den[0];
    }
}

// s = x - xp, y = g - gp; partials of y.s and y.y
__global__ void fit_update_pair_kernel(float *__restrict__ s, float *__restrict__ y, const float *__restrict__ x,
                                       const float *__restrict__ xp, const float *__restrict__ g,
                                       const float *__restrict__ gp, int64_t n, double *__restrict__ partial)
{
    __shared__ double s_red[32];
    double ays = 0.0, ayy = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float sv = x[e] - xp[e];
        const float yv = g[e] - gp[e];
        s[e] = sv;
        y[e] = yv;
        ays += (double)yv * (double)sv;
        ayy += (double)yv * (double)yv;
    }
    const double t0 = fit_block_sum(ays, s_red);
    const double t1 = fit_block_sum(ayy, s_red);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = t0;
        partial[FIT_BLOCKS + blockIdx.x] = t1;
    }
}

// two-loop building block:  d = (INIT ? -g : d + sign*coef[0]*v) * (num ? num[0]/den[0] : 1);  partial(u . d)
template <bool INIT>
__global__ void fit_axpy_dot_kernel(float *__restrict__ d, const float *__restrict__ g_or_v,
                                    const double *__restrict__ coef, float sign, const double *__restrict__ num,
                                    const double *__restrict__ den, const float *__restrict__ u, int64_t n,
                                    double *__restrict__ partial)
{
    __shared__ double s_red[32];
    const float c = INIT ? 0.f : sign * (float)coef[0];
    const float gamma = num != nullptr ? (float)(num[0] / den[0]) : 1.f;
    double acc = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float dv = INIT ? -g_or_v[e] : fmaf(c, g_or_v[e], d[e]);
        dv *= gamma;
        d[e] = dv;
        if (u != nullptr) acc += (double)u[e] * (double)dv;
    }
    if (u != nullptr) {
        const double tot = fit_block_sum(acc, s_red);
        if (threadIdx.x == 0) partial[blockIdx.x] = tot;
    }
}

// ---- host-side line search (More & Thuente) ---------------------------------------------------------
struct MtState {
    double x, fx, dx, y, fy, dy;
    bool brackt;
};

static double cubic_min(double u, double fu, double du, double v, double fv, double dv)
{
    const double d = v - u;
    const double theta = (fu - fv) * 3.0 / d + du + dv;
    const double s = std::max(std::fabs(theta), std::max(std::fabs(du), std::fabs(dv)));
    const double a = theta / s;
    double gamma = s * std::sqrt(std::max(0.0, a * a - (du / s) * (dv / s)));
    if (v < u) gamma = -gamma;
    const double p = gamma - du + theta;
    const double q = gamma - du + gamma + dv;
    return u + (p / q) * d;
}

static double cubic_min2(double u, double fu, double du, double v, double fv, double dv, double xmin, double xmax)
{
    const double d = v - u;
    const double theta = (fu - fv) * 3.0 / d + du + dv;
    const double s = std::max(std::fabs(theta), std::max(std::fabs(du), std::fabs(dv)));
    const double a = theta / s;
    double gamma = s * std::sqrt(std::max(0.0, a * a - (du / s) * (dv / s)));
    if (u < v) gamma = -gamma;
    const double p = gamma - dv + theta;
    const double q = gamma - dv + gamma + du;
    const double r = p / q;
    if (r < 0.0 && gamma != 0.0) return v - r * d;
    if (a < 0) return xmax;
    return xmin;
}

static double quad_min(double u, double fu, double du, double v, double fv)
{
    const double a = v - u;
    return u + du / ((fu - fv) / a + du) / 2.0 * a;
}

static double quad_min2(double u, double du, double v, double dv)
{
    const double a = u - v;
    return v + dv / (dv - du) * a;
}

// safeguarded trial-value update (More & Thuente sec. 4); returns true on an inconsistent interval
static bool update_trial_interval(MtState &st, double &t, double ft, double dt, double tmin, double tmax)
{
    double x = st.x, fx = st.fx, dx = st.dx, y = st.y, fy = st.fy, dy = st.dy;
    bool brackt = st.brackt;
    const bool dsign = dx != 0.0 ? (dt * (dx / std::fabs(dx)) < 0.0) : (dt < 0.0);
    if (brackt) {
        if (t <= std::min(x, y) || std::max(x, y) <= t) return true;
        if (0.0 <= dx * (t - x)) return true;
        if (tmax < tmin) return true;
    }
    bool bound;
    double newt;
    if (fx < ft) {
        brackt = true;
        bound = true;
        const double mc = cubic_min(x, fx, dx, t, ft, dt), mq = quad_min(x, fx, dx, t, ft);
        newt = std::fabs(mc - x) < std::fabs(mq - x) ? mc : mc + 0.5 * (mq - mc);
    } else if (dsign) {
        brackt = true;
        bound = false;
        const double mc = cubic_min(x, fx, dx, t, ft, dt), mq = quad_min2(x, dx, t, dt);
        newt = std::fabs(mc - t) > std::fabs(mq - t) ? mc : mq;
    } else if (std::fabs(dt) < std::fabs(dx)) {
        bound = true;
        const double mc = cubic_min2(x, fx, dx, t, ft, dt, tmin, tmax), mq = quad_min2(x, dx, t, dt);
        if (brackt) newt = std::fabs(t - mc) < std::fabs(t - mq) ? mc : mq;
        else newt = std::fabs(t - mc) > std::fabs(t - mq) ? mc : mq;
    } else {
        bound = false;
        if (brackt) newt = cubic_min(t, ft, dt, y, fy, dy);
        else if (x < t) newt = tmax;
        else newt = tmin;
    }
    if (fx < ft) {
        y = t; fy = ft; dy = dt;
    } else {
        if (dsign) { y = x; fy = fx; dy = dx; }
        x = t; fx = ft; dx = dt;
    }
    newt = std::min(tmax, std::max(tmin, newt));
    if (brackt && bound) {
        const double mq = x + 0.66 * (y - x);
        if (x < y) { if (mq < newt) newt = mq; }
        else { if (newt < mq) newt = mq; }
    }
    st.x = x; st.fx = fx; st.dx = dx; st.y = y; st.fy = fy; st.dy = dy; st.brackt = brackt;
    t = newt;
    return false;
}

// ---- the fit workspace (owned by the handle) -----------------------------------------------------------
struct FitWork {
    int64_t n = 0, stride = 0;
    int m = 0;
    float *x[2] = {nullptr, nullptr};   // current / trial parameters (ping-pong)
    float *g[2] = {nullptr, nullptr};   // gradients, each with 4 trailing floats for the packed -loglk
    float *d = nullptr;
    float *S = nullptr, *Y = nullptr;   // m x stride
    double *sc = nullptr;               // device scalars (SC_*)
    double *partial = nullptr;          // FIT_NRED * FIT_BLOCKS
    double *fx_data = nullptr;          // [2] data-term -loglk written by evc_plm_eval_data
    double *h_sc = nullptr;             // pinned host copy of sc[0..8)
};

void fit_work_free(FitWork *w)
{
    if (!w) return;
    for (int k = 0; k < 2; k++) { cudaFree(w->x[k]); cudaFree(w->g[k]); }
    cudaFree(w->d); cudaFree(w->S); cudaFree(w->Y); cudaFree(w->sc); cudaFree(w->partial); cudaFree(w->fx_data);
    if (w->h_sc) cudaFreeHost(w->h_sc);
    delete w;
}

static FitWork *fit_work_create(int64_t n, int m)
{
    FitWork *w = new (std::nothrow) FitWork();
    if (!w) return nullptr;
    w->n = n;
    w->m = m;
    w->stride = round_up(n + 4, 64);
    const size_t vb = (size_t)w->stride * sizeof(float);
    bool ok = true;
    for (int k = 0; k < 2 && ok; k++)
        ok = cudaMalloc(&w->x[k], vb) == cudaSuccess && cudaMalloc(&w->g[k], vb) == cudaSuccess;
    ok = ok && cudaMalloc(&w->d, vb) == cudaSuccess && cudaMalloc(&w->S, vb * m) == cudaSuccess &&
         cudaMalloc(&w->Y, vb * m) == cudaSuccess && cudaMalloc(&w->sc, SC_COUNT * sizeof(double)) == cudaSuccess &&
         cudaMalloc(&w->partial, (size_t)FIT_NRED * FIT_BLOCKS * sizeof(double)) == cudaSuccess &&
         cudaMalloc(&w->fx_data, 2 * sizeof(double)) == cudaSuccess &&
         cudaMallocHost(&w->h_sc, 8 * sizeof(double)) == cudaSuccess;
    if (!ok) {
        set_error(std::string("evc_plm_fit: workspace allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        fit_work_free(w);
        return nullptr;
    }
    cudaMemset(w->sc, 0, SC_COUNT * sizeof(double));
    return w;
}

struct FitCtx {
    evc_plm_t *h;
    FitWork *w;
    const evc_fit_params_t *p;
    evc_allreduce_cb ar;
    void *ar_user;
    cudaStream_t st;
    int evals = 0;
};

// objective + gradient at w->x[which] into w->g[which]; dvec (may be null) gives g.d.  Host scalars in w->h_sc.
static int fit_evaluate(FitCtx &c, int which, const float *dvec)
{
    FitWork *w = c.w;
    float *x = w->x[which], *g = w->g[which];
    if (evc_plm_eval_data(c.h, x, g, w->fx_data, c.st)) return 1;
    const float *limbs = nullptr;
    if (c.ar) {
        fit_pack_fx_kernel<<<1, 1, 0, c.st>>>(w->fx_data, g + w->n);
        EVC_KERNEL_CHECK();
        if (c.ar(c.ar_user, g, w->n + 4, c.st)) { set_error("evc_plm_fit: all-reduce callback failed"); return 1; }
        limbs = g + w->n;
    }
    const int64_t nh = (int64_t)c.h->g.L * c.h->g.q;
    fit_reg_dots_kernel<<<FIT_BLOCKS, FIT_THREADS, 0, c.st>>>(x, g, dvec, w->n, nh, c.p->lambda_h, c.p->lambda_J,
                                                              w->partial);
    EVC_KERNEL_CHECK();
    fit_reg_final_kernel<<<1, 1024, 0, c.st>>>(w->partial, w->fx_data, limbs, w->sc);
    EVC_KERNEL_CHECK();
    EVC_CUDA(cudaMemcpyAsync(w->h_sc, w->sc, 8 * sizeof(double), cudaMemcpyDeviceToHost, c.st));
    EVC_CUDA(cudaStreamSynchronize(c.st));
    c.evals++;
    return 0;
}

// d = -H g by the two-loop recursion over the `bound` newest pairs (ring of m, `end` = next slot to write)
static int fit_direction(FitCtx &c, int cur, int bound, int end)
{
    FitWork *w = c.w;
    const int m = w->m;
    const int64_t n = w->n, ld = w->stride;
    float *d = w->d;
    const float *g = w->g[cur];
    double *ys = w->sc + SC_YS, *alpha = w->sc + SC_ALPHA, *coef = w->sc + SC_COEF, *yy = w->sc + SC_YY;
    if (bound == 0) {
        fit_axpy_dot_kernel<true><<<FIT_BLOCKS, FIT_THREADS, 0, c.st>>>(d, g, nullptr, 0.f, nullptr, nullptr, nullptr, n,
                                                                       w->partial);
        EVC_KERNEL_CHECK();
        return 0;
    }
    const int newest = (end + m - 1) % m;
    int j = newest;
    // d = -g; alpha_newest = (s_newest . d) / ys
    fit_axpy_dot_kernel<true><<<FIT_BLOCKS, FIT_THREADS, 0, c.st>>>(d, g, nullptr, 0.f, nullptr, nullptr,
                                                                   w->S + (int64_t)j * ld, n, w->partial);
    EVC_KERNEL_CHECK();
    fit_scalar_final_kernel<<<1, 1024, 0, c.st>>>(w->partial, 1, ys + j, nullptr, alpha + j);
    EVC_KERNEL_CHECK();
    for (int it = 0; it < bound; it++) {
        const bool last = it == bound - 1;
        if (!last) {
            const int jn = (j + m - 1) % m;
            // d -= alpha_j y_j; alpha_jn = (s_jn . d) / ys_jn
            fit_axpy_dot_kernel<false><<<FIT_BLOCKS, FIT_THREADS, 0, c.st>>>(d, w->Y + (int64_t)j * ld, alpha + j, -1.f,
                                                                            nullptr, nullptr, w->S + (int64_t)jn * ld, n,
                                                                            w->partial);
            EVC_KERNEL_CHECK();
            fit_scalar_final_kernel<<<1, 1024, 0, c.st>>>(w->partial, 1, ys + jn, nullptr, alpha + jn);
            EVC_KERNEL_CHECK();
            j = jn;
        } else {
            // oldest pair: d = (d - alpha_j y_j) * ys_newest / yy_newest; coef = alpha_j - (y_j . d) / ys_j
            fit_axpy_dot_kernel<false><<<FIT_BLOCKS, FIT_THREADS, 0, c.st>>>(d, w->Y + (int64_t)j * ld, alpha + j, -1.f,
                                                                            ys + newest, yy, w->Y + (int64_t)j * ld, n,
                                                                            w->partial);
            EVC_KERNEL_CHECK();
            fit_scalar_final_kernel<<<1, 1024, 0, c.st>>>(w->partial, 2, ys + j, alpha + j, coef);
            EVC_KERNEL_CHECK();
        }
    }
    for (int it = 0; it < bound; it++) {
        const bool last = it == bound - 1;
        const int jn = (j + 1) % m;
        // d += coef s_j; coef' = alpha_jn - (y_jn . d) / ys_jn
        fit_axpy_dot_kernel<false><<<FIT_BLOCKS, FIT_THREADS, 0, c.st>>>(d, w->S + (int64_t)j * ld, coef, 1.f, nullptr,
                                                                        nullptr, last ? nullptr : w->Y + (int64_t)jn * ld,
                                                                        n, w->partial);
        EVC_KERNEL_CHECK();
        if (!last) {
            fit_scalar_final_kernel<<<1, 1024, 0, c.st>>>(w->partial, 2, ys + jn, alpha + jn, coef);
            EVC_KERNEL_CHECK();
        }
        j = jn;
    }
    return 0;
}

}  // namespace evc

using namespace evc;

extern "C" {

void evc_fit_default_params(evc_fit_params_t *p)
{
    if (!p) return;
    p->max_iterations = 0;
    p->m = 6;
    p->epsilon = 1e-3f;
    p->lambda_h = 0.01f;
    p->lambda_J = 100.f;
    p->max_linesearch = 40;
    p->min_step = 1e-20;
    p->max_step = 1e20;
    p->ftol = 1e-4;
    p->gtol = 0.9;
    p->xtol = 1e-7;
    p->precision_schedule = 0;
    p->switch_factor = 10.f;
}

int evc_plm_fit(evc_plm_t *h, float *d_x, const evc_fit_params_t *p, evc_allreduce_cb allreduce, void *allreduce_user,
                evc_progress_cb progress, void *progress_user, evc_fit_result_t *res, void *stream)
{
    if (!h || !d_x || !p || !res) { set_error("evc_plm_fit: null pointer"); return 1; }
    if (p->m < 1 || p->m > 32) { set_error("evc_plm_fit: history m must be in 1..32"); return 1; }
    const int64_t n = evc_plm_num_params(h);
    EVC_CUDA(cudaSetDevice(h->device));
    if (h->fit && h->fit->m != p->m) {
        fit_work_free(h->fit);
        h->fit = nullptr;
    }
    if (!h->fit) h->fit = fit_work_create(n, p->m);
    FitWork *w = h->fit;
    if (!w) return 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FitCtx c{h, w, p, allreduce, allreduce_user, st};
    const auto t_begin = std::chrono::steady_clock::now();
    const size_t nb = (size_t)n * sizeof(float);
    int cur = 0;
    int k = 0, switched_at = -1;
    bool low = false;
    if (p->precision_schedule == 1) {
        if (evc_plm_set_precision(h, 1)) return 1;
        low = true;
    }
    EVC_CUDA(cudaMemcpyAsync(w->x[cur], d_x, nb, cudaMemcpyDeviceToDevice, st));
    if (fit_evaluate(c, cur, nullptr)) return 1;
    double fx = w->h_sc[SC_FX], nll = w->h_sc[SC_NLL];
    double xnorm = std::sqrt(w->h_sc[SC_XXH] + w->h_sc[SC_XXJ]), gnorm = std::sqrt(w->h_sc[SC_GG]);
    auto finish = [&](int stat) {
        res->status = stat;
        res->iterations = k;
        res->evaluations = c.evals;
        res->switched_at = switched_at;
        res->fx = fx;
        res->negloglk = nll;
        res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
        if (cudaMemcpyAsync(d_x, w->x[cur], nb, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("evc_plm_fit: copying the result failed");
            return 1;
        }
        return 0;
    };
    // leave the bf16x1 mode: hi+lo products from here on, objective re-evaluated at x[cur], history dropped
    // (a stored pair would mix gradients of two precisions), restart from steepest descent
    int hist = 0, end = 0;
    double step = 0.0;
    auto switch_to_high = [&](int at) -> int {
        if (evc_plm_set_precision(h, 0)) return 1;
        low = false;
        switched_at = at;
        if (fit_evaluate(c, cur, nullptr)) return 1;
        fx = w->h_sc[SC_FX];
        nll = w->h_sc[SC_NLL];
        xnorm = std::sqrt(w->h_sc[SC_XXH] + w->h_sc[SC_XXJ]);
        gnorm = std::sqrt(w->h_sc[SC_GG]);
        hist = 0;
        end = 0;
        if (fit_direction(c, cur, 0, 0)) return 1;
        step = 1.0 / gnorm;
        return 0;
    };
    if (gnorm / std::max(1.0, xnorm) <= p->epsilon) {
        if (!low) return finish(EVC_LBFGS_ALREADY_MINIMIZED);
        if (switch_to_high(0)) return 1;
        if (gnorm / std::max(1.0, xnorm) <= p->epsilon) return finish(EVC_LBFGS_ALREADY_MINIMIZED);
    }
    if (fit_direction(c, cur, 0, 0)) return 1;
    step = 1.0 / gnorm;
    k = 1;
    for (;;) {
        // ---- line search along d from x[cur] ----
        fit_dot_kernel<<<FIT_BLOCKS, FIT_THREADS, 0, st>>>(w->g[cur], w->d, n, w->partial);
        EVC_KERNEL_CHECK();
        fit_scalar_final_kernel<<<1, 1024, 0, st>>>(w->partial, 0, nullptr, nullptr, w->sc + SC_DG);
        EVC_KERNEL_CHECK();
        EVC_CUDA(cudaMemcpyAsync(w->h_sc, w->sc, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
        EVC_CUDA(cudaStreamSynchronize(st));
        const double finit = fx, dginit = w->h_sc[SC_DG];
        const int trial = cur ^ 1;
        int ls_status = 0;      // 0 = the line search converged (strong Wolfe conditions hold at `step`)
        int count = 0;
        double f = finit;
        if (step <= 0.0) ls_status = EVC_LBFGSERR_INVALIDPARAMETERS;
        else if (dginit > 0.0) ls_status = EVC_LBFGSERR_INCREASEGRADIENT;
        else {
            MtState ms{0.0, finit, dginit, 0.0, finit, dginit, false};
            bool stage1 = true, uinfo = false;
            const double dgtest = p->ftol * dginit;
            double width = p->max_step - p->min_step, prev_width = 2.0 * width;
            for (;;) {
                double stmin, stmax;
                if (ms.brackt) { stmin = std::min(ms.x, ms.y); stmax = std::max(ms.x, ms.y); }
                else { stmin = ms.x; stmax = step + 4.0 * (step - ms.x); }
                step = std::min(p->max_step, std::max(p->min_step, step));
                if ((ms.brackt && ((step <= stmin || stmax <= step) || p->max_linesearch <= count + 1 || uinfo)) ||
                    (ms.brackt && (stmax - stmin <= p->xtol * stmax)))
                    step = ms.x;
                fit_step_kernel<<<FIT_BLOCKS, FIT_THREADS, 0, st>>>(w->x[trial], w->x[cur], w->d, (float)step, n);
                EVC_KERNEL_CHECK();
                if (fit_evaluate(c, trial, w->d)) return 1;
                f = w->h_sc[SC_FX];
                const double dg = w->h_sc[SC_DG];
                const double ftest1 = finit + step * dgtest;
                count++;
                if (ms.brackt && ((step <= stmin || stmax <= step) || uinfo)) { ls_status = EVC_LBFGSERR_ROUNDING_ERROR; break; }
                if (step == p->max_step && f <= ftest1 && dg <= dgtest) { ls_status = EVC_LBFGSERR_MAXIMUMSTEP; break; }
                if (step == p->min_step && (ftest1 < f || dgtest <= dg)) { ls_status = EVC_LBFGSERR_MINIMUMSTEP; break; }
                if (ms.brackt && (stmax - stmin) <= p->xtol * stmax) { ls_status = EVC_LBFGSERR_WIDTHTOOSMALL; break; }
                if (p->max_linesearch <= count) { ls_status = EVC_LBFGSERR_MAXIMUMLINESEARCH; break; }
                if (f <= ftest1 && std::fabs(dg) <= p->gtol * (-dginit)) break;     // accept
                if (stage1 && f <= ftest1 && std::min(p->ftol, p->gtol) * dginit <= dg) stage1 = false;
                if (stage1 && ftest1 < f && f <= ms.fx) {
                    MtState m2{ms.x, ms.fx - ms.x * dgtest, ms.dx - dgtest, ms.y, ms.fy - ms.y * dgtest, ms.dy - dgtest,
                               ms.brackt};
                    uinfo = update_trial_interval(m2, step, f - step * dgtest, dg - dgtest, stmin, stmax);
                    ms = MtState{m2.x, m2.fx + m2.x * dgtest, m2.dx + dgtest, m2.y, m2.fy + m2.y * dgtest,
                                 m2.dy + dgtest, m2.brackt};
                } else {
                    uinfo = update_trial_interval(ms, step, f, dg, stmin, stmax);
                }
                if (ms.brackt) {
                    if (0.66 * prev_width <= std::fabs(ms.y - ms.x)) step = ms.x + 0.5 * (ms.y - ms.x);
                    prev_width = width;
                    width = std::fabs(ms.y - ms.x);
                }
            }
        }
        if (ls_status != 0) {
            if (low) {
                // the bf16x1 gradient is no longer good enough for the line search: finish in the hi+lo mode
                if (switch_to_high(k)) return 1;
                continue;
            }
            k = k - 1;
            return finish(ls_status);     // x[cur], g[cur] are the last accepted point
        }
        // ---- accepted: x[trial] is the new iterate ----
        const int prev = cur;
        cur = trial;
        fx = f;
        nll = w->h_sc[SC_NLL];
        xnorm = std::sqrt(w->h_sc[SC_XXH] + w->h_sc[SC_XXJ]);
        gnorm = std::sqrt(w->h_sc[SC_GG]);
        if (progress && progress(progress_user, k, fx, xnorm, gnorm, step, count, nll, std::sqrt(w->h_sc[SC_XXH]),
                                 std::sqrt(w->h_sc[SC_XXJ])))
            return finish(EVC_LBFGSERR_CANCELED);
        if (low && gnorm / std::max(1.0, xnorm) <= (double)p->switch_factor * p->epsilon) {
            if (switch_to_high(k)) return 1;
            if (gnorm / std::max(1.0, xnorm) <= p->epsilon) return finish(EVC_LBFGS_SUCCESS);
            if (p->max_iterations != 0 && p->max_iterations < k + 1) return finish(EVC_LBFGSERR_MAXIMUMITERATION);
            k++;
            continue;
        }
        if (gnorm / std::max(1.0, xnorm) <= p->epsilon) return finish(EVC_LBFGS_SUCCESS);
        if (p->max_iterations != 0 && p->max_iterations < k + 1) return finish(EVC_LBFGSERR_MAXIMUMITERATION);
        // correction pair into slot `end`
        fit_update_pair_kernel<<<FIT_BLOCKS, FIT_THREADS, 0, st>>>(w->S + (int64_t)end * w->stride,
                                                                  w->Y + (int64_t)end * w->stride, w->x[cur], w->x[prev],
                                                                  w->g[cur], w->g[prev], n, w->partial);
        EVC_KERNEL_CHECK();
        fit_scalar_final_kernel<<<1, 1024, 0, st>>>(w->partial, 0, nullptr, nullptr, w->sc + SC_YS + end);
        EVC_KERNEL_CHECK();
        fit_scalar_final_kernel<<<1, 1024, 0, st>>>(w->partial + FIT_BLOCKS, 0, nullptr, nullptr, w->sc + SC_YY);
        EVC_KERNEL_CHECK();
        hist = std::min(p->m, hist + 1);
        end = (end + 1) % p->m;
        k++;
        if (fit_direction(c, cur, hist, end)) return 1;
        step = 1.0;
    }
}

}  // extern "C"

// One collective per evaluation outside evc_plm_fit as well (bench.py / the Python driver): -loglk rides behind
// the gradient as exact fixed-point limbs (see fit_pack_fx_kernel).
namespace evc {
__global__ void fit_unpack_fx_kernel(const float *__restrict__ limbs, double *__restrict__ fx)
{
    const long long q = (long long)limbs[0] + ((long long)limbs[1] << FX_LIMB_BITS) +
                        ((long long)limbs[2]) * (1ll << (2 * FX_LIMB_BITS));
    fx[0] = (double)q / FX_SCALE;
}
}  // namespace evc

extern "C" {
int evc_plm_pack_fx(const double *d_fx, float *d_limbs, void *stream)
{
    if (!d_fx || !d_limbs) { set_error("evc_plm_pack_fx: null pointer"); return 1; }
    fit_pack_fx_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_fx, d_limbs);
    EVC_KERNEL_CHECK();
    return 0;
}
int evc_plm_unpack_fx(const float *d_limbs, double *d_fx, void *stream)
{
    if (!d_fx || !d_limbs) { set_error("evc_plm_unpack_fx: null pointer"); return 1; }
    fit_unpack_fx_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_limbs, d_fx);
    EVC_KERNEL_CHECK();
    return 0;
}
}  // extern "C"
