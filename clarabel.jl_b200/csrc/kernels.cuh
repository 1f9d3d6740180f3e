// Device kernels for the B200 KKT path (sm_100a).  FP64 throughout (the reference is Float64
// end-to-end and its IR tolerance is 1e-13 — src/settings.jl:127-132).
//
//   G1  cone -> K value update            k_hs_diag, k_hs_soc_dense, k_soc_expansion
//   G2  PSD skron                         k_psd_rrt, k_psd_skron
//   G3  static regularisation             k_diag_absmax, k_compute_eps, k_shift_diag
//   G4  small-front LDL' (shared memory)  k_factor_small
//   G5  large-front blocked LDL'          k_panel_large, k_update_large
//   G6  extend-add                        (fused in k_factor_small) / k_assemble_large
//   G7  multifrontal triangular solves    k_fwd, k_bwd, k_pack_perm, k_unpack_perm
//   G8  symmetric SpMV residual + norm    k_residual
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cb200 {

struct DevSym {                 // device copies of the Symbolic arrays
    const int32_t* sn_first;    // [nsuper+1]
    const int64_t* rows_ptr;    // [nsuper+1]
    const int32_t* rows;
    const int32_t* rel;
    const int32_t* child_ptr;
    const int32_t* child_list;
    const int64_t* panel_off;
    const int64_t* upd_off;
    const int8_t*  dsign;       // [N] permuted pivot signs
};

struct RegParams { double eps, delta; int enable; };

// ------------------------------------------------------------------ G4 small fronts
// One CTA per front; the whole front lives in shared memory.  Fuses: extend-add of the children's
// update blocks (G6), right-looking LDL' of the ns pivot columns with the reference's sign-based
// dynamic regularisation (QDLDL semantics: if D[k]*sign < eps then D[k] = delta*sign), and the
// write-back of the L panel and of this front's update block.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_factor_small(DevSym S, const int32_t* __restrict__ batch, double* __restrict__ Lst,
               double* __restrict__ Ust, double* __restrict__ D, double* __restrict__ Dinv,
               RegParams rp, unsigned int* __restrict__ nreg) {
    extern __shared__ double F[];
    const int s = batch[blockIdx.x];
    const int f = S.sn_first[s];
    const int ns = S.sn_first[s + 1] - f;
    const int nr = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
    const int nf = ns + nr;
    const int tid = threadIdx.x;
    double* Lp = Lst + S.panel_off[s];
    // 1. panel -> first ns columns (identical layout), zero the trailing block
    for (int i = tid; i < nf * ns; i += THREADS) F[i] = Lp[i];
    for (int i = nf * ns + tid; i < nf * nf; i += THREADS) F[i] = 0.0;
    __syncthreads();
    // 2. extend-add children (sequential over children => deterministic sums)
    for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; ++q) {
        const int c = S.child_list[q];
        const int64_t rp0 = S.rows_ptr[c];
        const int nrc = (int)(S.rows_ptr[c + 1] - rp0);
        const int32_t* relc = S.rel + rp0;
        const double* Uc = Ust + S.upd_off[c];
        if (THREADS >= 64) {
            const int lane = tid & 31, wid = tid >> 5;
            for (int j = wid; j < nrc; j += THREADS / 32) {
                const int dj = relc[j] * nf;
                for (int i = j + lane; i < nrc; i += 32) F[relc[i] + dj] += Uc[i + (int64_t)j * nrc];
            }
        } else {
            for (int e = tid; e < nrc * nrc; e += THREADS) {
                const int i = e % nrc, j = e / nrc;
                if (i >= j) F[relc[i] + relc[j] * nf] += Uc[e];
            }
        }
        __syncthreads();
    }
    // 3. right-looking LDL' on the pivot columns
    const int lane = tid & 31, wid = tid >> 5;
    constexpr int NW = THREADS / 32;
    for (int k = 0; k < ns; ++k) {
        double d = F[k + k * nf];
        const double sg = (double)S.dsign[f + k];
        bool reg = false;
        if (rp.enable && d * sg < rp.eps) { d = rp.delta * sg; reg = true; }
        const double dinv = 1.0 / d;
        const double* ck = F + k * nf;
        for (int j = k + 1 + wid; j < nf; j += NW) {
            const double wj = ck[j] * dinv;
            double* cj = F + j * nf;
            for (int i = j + lane; i < nf; i += 32) cj[i] -= ck[i] * wj;
        }
        __syncthreads();
        for (int i = k + 1 + tid; i < nf; i += THREADS) F[i + k * nf] *= dinv;
        if (tid == 0) {
            F[k + k * nf] = d; D[f + k] = d; Dinv[f + k] = dinv;
            if (reg) atomicAdd(nreg, 1u);
        }
        __syncthreads();
    }
    // 4. write back
    for (int i = tid; i < nf * ns; i += THREADS) Lp[i] = F[i];
    if (nr > 0) {
        double* Us = Ust + S.upd_off[s];
        for (int j = wid; j < nr; j += NW) {
            const double* cj = F + (ns + j) * nf + ns;
            for (int i = j + lane; i < nr; i += 32) Us[i + (int64_t)j * nr] = cj[i];
        }
    }
}

// ------------------------------------------------------------------ G6 assembly for large fronts
// grid (chunks, nbatch).  Chunk c owns destination columns [c*CW, (c+1)*CW) of the parent front,
// so different CTAs never write the same entry; children are processed in order.
constexpr int ASM_CW = 32;
__global__ void __launch_bounds__(256)
k_assemble_large(DevSym S, const int32_t* __restrict__ batch, double* __restrict__ Lst,
                 double* __restrict__ Ust) {
    const int s = batch[blockIdx.y];
    const int f = S.sn_first[s];
    const int ns = S.sn_first[s + 1] - f;
    const int nr = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
    const int nf = ns + nr;
    const int c0 = blockIdx.x * ASM_CW, c1 = min(nf, c0 + ASM_CW);
    if (c0 >= nf) return;
    double* Lp = Lst + S.panel_off[s];
    double* Us = Ust + S.upd_off[s];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; ++q) {
        const int c = S.child_list[q];
        const int64_t rp0 = S.rows_ptr[c];
        const int nrc = (int)(S.rows_ptr[c + 1] - rp0);
        const int32_t* relc = S.rel + rp0;
        const double* Uc = Ust + S.upd_off[c];
        // child columns whose destination falls into [c0, c1): rel is increasing -> binary search
        int lo = 0, hi = nrc;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < c0) lo = mid + 1; else hi = mid; }
        const int jb = lo;
        hi = nrc;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < c1) lo = mid + 1; else hi = mid; }
        const int je = lo;
        for (int j = jb + wid; j < je; j += 8) {
            const int dj = relc[j];
            double* dst = dj < ns ? Lp + (int64_t)dj * nf : Us + (int64_t)(dj - ns) * nr - ns;
            const double* src = Uc + (int64_t)j * nrc;
            for (int i = j + lane; i < nrc; i += 32) dst[relc[i]] += src[i];
        }
        __syncthreads();   // (only orders this CTA's own adds; kept for clarity)
    }
}

// ------------------------------------------------------------------ G5 large fronts, blocked
constexpr int LNB = 32;        // pivot block width
constexpr int LTR = 128;       // rows per CTA in the panel kernel

// Front-local element address: column g < ns lives in the panel, otherwise in the update block.
__device__ __forceinline__ double* front_col(double* Lp, double* Us, int ns, int nr, int nf, int g) {
    return g < ns ? Lp + (int64_t)g * nf : Us + (int64_t)(g - ns) * nr - ns;
}

// grid (row tiles, nbatch), 128 threads.  Every CTA factors the nb x nb diagonal block in shared
// memory (redundantly: no inter-CTA dependency), then solves its rows against it.
__global__ void __launch_bounds__(LTR)
k_panel_large(DevSym S, const int32_t* __restrict__ batch, int kb, double* __restrict__ Lst,
              double* __restrict__ Wst, const int64_t* __restrict__ woff,
              double* __restrict__ D, double* __restrict__ Dinv, RegParams rp,
              unsigned int* __restrict__ nreg) {
    __shared__ double A[LNB][LNB + 1];
    __shared__ double dv[LNB], dinvs[LNB];
    const int s = batch[blockIdx.y];
    const int f = S.sn_first[s];
    const int ns = S.sn_first[s + 1] - f;
    if (kb >= ns) return;
    const int nr = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
    const int nf = ns + nr;
    const int nb = min(LNB, ns - kb);
    const int r0 = kb + nb + blockIdx.x * LTR;
    if (r0 >= nf && blockIdx.x != 0) return;
    double* Lp = Lst + S.panel_off[s];
    const int tid = threadIdx.x;
    // load diagonal block (lower part)
    for (int e = tid; e < nb * nb; e += LTR) {
        const int i = e % nb, j = e / nb;
        A[i][j] = (i >= j) ? Lp[(int64_t)(kb + j) * nf + kb + i] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        double d = A[k][k];
        const double sg = (double)S.dsign[f + kb + k];
        bool reg = false;
        if (rp.enable && d * sg < rp.eps) { d = rp.delta * sg; reg = true; }
        const double dinv = 1.0 / d;
        __syncthreads();
        // trailing update inside the block: thread (i,j) pairs
        for (int e = tid; e < (nb - k - 1) * (nb - k - 1); e += LTR) {
            const int i = k + 1 + e % (nb - k - 1), j = k + 1 + e / (nb - k - 1);
            if (i >= j) A[i][j] -= A[i][k] * A[j][k] * dinv;
        }
        __syncthreads();
        for (int i = k + 1 + tid; i < nb; i += LTR) A[i][k] *= dinv;
        if (tid == 0) {
            dv[k] = d; dinvs[k] = dinv;
            if (blockIdx.x == 0 && reg) atomicAdd(nreg, 1u);
        }
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        // The factored diagonal block cannot be written in place here: sibling CTAs of this
        // front may still be loading the unfactored block.  Park it in the (otherwise unused)
        // rows [kb, kb+nb) of the W workspace; k_update_large copies it into the panel.
        double* Wd = Wst + woff[blockIdx.y];
        for (int e = tid; e < nb * nb; e += LTR) {
            const int i = e % nb, j = e / nb;
            if (i > j) Wd[(int64_t)j * nf + kb + i] = A[i][j];
            else if (i == j) Wd[(int64_t)j * nf + kb + i] = dv[j];
        }
        for (int k = tid; k < nb; k += LTR) { D[f + kb + k] = dv[k]; Dinv[f + kb + k] = dinvs[k]; }
    }
    // rows below the diagonal block: x * L11' = a ; L = x * D^-1 ; W = x
    const int r = r0 + tid;
    if (r < nf) {
        double x[LNB];
#pragma unroll
        for (int j = 0; j < LNB; ++j) x[j] = (j < nb) ? Lp[(int64_t)(kb + j) * nf + r] : 0.0;
#pragma unroll
        for (int j = 0; j < LNB; ++j) {
            if (j < nb) {
                double v = x[j];
#pragma unroll
                for (int l = 0; l < LNB; ++l) if (l < j) v -= x[l] * A[j][l];
                x[j] = v;
            }
        }
        double* W = Wst + woff[blockIdx.y];
#pragma unroll
        for (int j = 0; j < LNB; ++j) {
            if (j < nb) {
                Lp[(int64_t)(kb + j) * nf + r] = x[j] * dinvs[j];
                W[(int64_t)j * nf + r] = x[j];
            }
        }
    }
}

// grid (tile pairs, nbatch), 256 threads; 64x64 tile of the trailing matrix, K = nb (<= 32):
// C[i][j] -= sum_k L[i][k] * W[j][k]   for i >= j (lower part), i, j in [kb+nb, nf).
constexpr int UT = 64;
__global__ void __launch_bounds__(256)
k_update_large(DevSym S, const int32_t* __restrict__ batch, int kb, double* __restrict__ Lst,
               double* __restrict__ Ust, const double* __restrict__ Wst,
               const int64_t* __restrict__ woff) {
    __shared__ double sL[LNB][UT + 1];
    __shared__ double sW[LNB][UT + 1];
    const int s = batch[blockIdx.y];
    const int f = S.sn_first[s];
    const int ns = S.sn_first[s + 1] - f;
    if (kb >= ns) return;
    const int nr = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
    const int nf = ns + nr;
    const int nb = min(LNB, ns - kb);
    const int t0 = kb + nb;
    const int nt = nf - t0;
    if (blockIdx.x == 0) {      // move the factored diagonal block from W into the panel
        const double* Wd = Wst + woff[blockIdx.y];
        double* Lpd = Lst + S.panel_off[s];
        for (int e = threadIdx.x; e < nb * nb; e += 256) {
            const int i = e % nb, j = e / nb;
            if (i >= j) Lpd[(int64_t)(kb + j) * nf + kb + i] = Wd[(int64_t)j * nf + kb + i];
        }
    }
    if (nt <= 0) return;
    // linear tile index -> (ti >= tj)
    const int T = (nt + UT - 1) / UT;
    int idx = blockIdx.x;
    if (idx >= T * (T + 1) / 2) return;
    int ti = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= idx) ++ti;
    while (ti * (ti + 1) / 2 > idx) --ti;
    const int tj = idx - ti * (ti + 1) / 2;
    const int i0 = t0 + ti * UT, j0 = t0 + tj * UT;
    double* Lp = Lst + S.panel_off[s];
    double* Us = Ust + S.upd_off[s];
    const double* W = Wst + woff[blockIdx.y];
    const int tid = threadIdx.x;
    for (int e = tid; e < UT * LNB; e += 256) {
        const int r = e % UT, k = e / UT;
        const int gi = i0 + r, gj = j0 + r;
        sL[k][r] = (k < nb && gi < nf) ? Lp[(int64_t)(kb + k) * nf + gi] : 0.0;
        sW[k][r] = (k < nb && gj < nf) ? W[(int64_t)k * nf + gj] : 0.0;
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;     // 16 x 16 threads, 4 x 4 micro-tile each
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
    for (int k = 0; k < LNB; ++k) {
        double l[4], w[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { l[a] = sL[k][tx + 16 * a]; w[a] = sW[k][ty + 16 * a]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += l[a] * w[b];
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int gj = j0 + ty + 16 * b;
        if (gj >= nf) continue;
        double* col = front_col(Lp, Us, ns, nr, nf, gj);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int gi = i0 + tx + 16 * a;
            if (gi < nf && gi >= gj) col[gi] -= acc[a][b];
        }
    }
}

// ------------------------------------------------------------------ scatter / regularisation
__global__ void k_scatter(const double* __restrict__ nz, const int64_t* __restrict__ amap,
                          int64_t n, double* __restrict__ Lst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) Lst[amap[i]] = nz[i];
}

__global__ void k_diag_absmax(const double* __restrict__ nz, const int64_t* __restrict__ didx,
                              int64_t n, unsigned long long* __restrict__ out) {
    double m = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, fabs(nz[didx[i]]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// eps = c + p * max|diag|   (_compute_regularizer, kktsolver_directldl.jl:297-310)
__global__ void k_compute_eps(const unsigned long long* maxbits, double c, double p, double* eps) {
    // separate rounding of the product and the sum, like the reference's scalar code
    *eps = __dadd_rn(c, __dmul_rn(p, __longlong_as_double((long long)*maxbits)));
}

// Lst[amap[didx[i]]] = nz[didx[i]] + sign_i * eps  (static regularisation, :266-273; the
// unshifted nz is kept for the refinement residuals, :283-291)
__global__ void k_shift_diag(const double* __restrict__ nz, const int64_t* __restrict__ didx,
                             const int64_t* __restrict__ amap, const int8_t* __restrict__ dsign_orig,
                             const double* __restrict__ eps, int64_t n, double* __restrict__ Lst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t k = didx[i];
        Lst[amap[k]] = nz[k] + (dsign_orig[i] > 0 ? *eps : -*eps);
    }
}

// ------------------------------------------------------------------ G7 triangular solves
__global__ void k_pack_perm(const double* __restrict__ b, const int32_t* __restrict__ perm,
                            int64_t n, double* __restrict__ y) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] = b[perm[i]];
}
__global__ void k_unpack_perm(const double* __restrict__ y, const int32_t* __restrict__ perm,
                              int64_t n, double* __restrict__ x) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[perm[i]] = y[i];
}

// Forward sweep, multifrontal form (deterministic, no atomics): per supernode
//   w = [y_s ; 0] + sum_children extend(u_c);  w <- L^{-1} w (unit lower, all nf rows);
//   y_s = w[0:ns] * (1: stored unscaled),  u_s = w[ns:]
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_fwd(DevSym S, const int32_t* __restrict__ batch, const double* __restrict__ Lst,
      double* __restrict__ y, double* __restrict__ uvec) {
    extern __shared__ double w[];
    const int s = batch[blockIdx.x];
    const int f = S.sn_first[s];
    const int ns = S.sn_first[s + 1] - f;
    const int64_t rp = S.rows_ptr[s];
    const int nr = (int)(S.rows_ptr[s + 1] - rp);
    const int nf = ns + nr;
    const int tid = threadIdx.x;
    for (int i = tid; i < nf; i += THREADS) w[i] = i < ns ? y[f + i] : 0.0;
    __syncthreads();
    for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; ++q) {
        const int c = S.child_list[q];
        const int64_t rp0 = S.rows_ptr[c];
        const int nrc = (int)(S.rows_ptr[c + 1] - rp0);
        for (int i = tid; i < nrc; i += THREADS) w[S.rel[rp0 + i]] += uvec[rp0 + i];
        __syncthreads();
    }
    const double* Lp = Lst + S.panel_off[s];
    for (int k = 0; k < ns; ++k) {
        const double yk = w[k];
        const double* ck = Lp + (int64_t)k * nf;
        for (int i = k + 1 + tid; i < nf; i += THREADS) w[i] -= ck[i] * yk;
        __syncthreads();
    }
    for (int i = tid; i < nf; i += THREADS) {
        if (i < ns) y[f + i] = w[i]; else uvec[rp + i - ns] = w[i];
    }
}

// Backward sweep: x_s = L11^{-T} (D^{-1} y_s - L21' x[R_s])
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_bwd(DevSym S, const int32_t* __restrict__ batch, const double* __restrict__ Lst,
      const double* __restrict__ Dinv, double* __restrict__ y) {
    extern __shared__ double w[];
    __shared__ double red[32];
    const int s = batch[blockIdx.x];
    const int f = S.sn_first[s];
    const int ns = S.sn_first[s + 1] - f;
    const int64_t rp = S.rows_ptr[s];
    const int nr = (int)(S.rows_ptr[s + 1] - rp);
    const int nf = ns + nr;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < nf; i += THREADS)
        w[i] = i < ns ? y[f + i] * Dinv[f + i] : y[S.rows[rp + i - ns]];
    __syncthreads();
    const double* Lp = Lst + S.panel_off[s];
    for (int k = ns - 1; k >= 0; --k) {
        const double* ck = Lp + (int64_t)k * nf;
        double acc = 0.0;
        for (int i = k + 1 + tid; i < nf; i += THREADS) acc += ck[i] * w[i];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (THREADS > 32) {
            if (lane == 0) red[wid] = acc;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int q = 0; q < THREADS / 32; ++q) t += red[q];
                w[k] -= t;
            }
            __syncthreads();
        } else {
            if (tid == 0) w[k] -= acc;
            __syncwarp();
        }
    }
    for (int i = tid; i < ns; i += THREADS) y[f + i] = w[i];
}

// ------------------------------------------------------------------ G8 residual e = b - K x
// K symmetric, stored as upper CSC (cp, ri, nz) plus the row-wise index of the same entries
// (tp, tc, tpos: entries (j, c > j) of row j, value nz[tpos]).  One warp per row.
__global__ void __launch_bounds__(256)
k_residual(int64_t N, const int64_t* __restrict__ cp, const int32_t* __restrict__ ri,
           const double* __restrict__ nz, const int64_t* __restrict__ tp,
           const int32_t* __restrict__ tc, const int64_t* __restrict__ tpos,
           const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ e,
           unsigned long long* __restrict__ norm_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    double r = 0.0;
    if (row < N) {
        double acc = 0.0;
        for (int64_t p = cp[row] + lane; p < cp[row + 1]; p += 32) acc += nz[p] * x[ri[p]];
        for (int64_t p = tp[row] + lane; p < tp[row + 1]; p += 32) acc += nz[tpos[p]] * x[tc[p]];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        r = b[row] - acc;
        if (lane == 0) e[row] = r;
    }
    double m = fabs(r);
    if (!(m == m)) m = __longlong_as_double(0x7ff0000000000000LL);   // NaN -> +inf
    __shared__ double sm[8];
    if (lane == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t = fmax(t, sm[q]);
        atomicMax(norm_bits, (unsigned long long)__double_as_longlong(t));
    }
}

__global__ void k_absmax(const double* __restrict__ v, int64_t n, unsigned long long* __restrict__ out) {
    double m = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double a = fabs(v[i]);
        if (!(a == a)) a = __longlong_as_double(0x7ff0000000000000LL);
        m = fmax(m, a);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

__global__ void k_axpy1(double* __restrict__ dx, const double* __restrict__ x, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dx[i] += x[i];
}

__global__ void k_build_rhs(const double* __restrict__ rx, const double* __restrict__ rz,
                            int64_t n, int64_t m, int64_t N, double* __restrict__ b) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) b[i] = i < n ? rx[i] : (i < n + m ? rz[i - n] : 0.0);
}

// update_values!/scale_values! on the device copy (inner boundary)
__global__ void k_update_values(double* __restrict__ nz, const int64_t* __restrict__ idx,
                                const double* __restrict__ v, int64_t n, int64_t base) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) nz[idx[i] - base] = v[i];
}
__global__ void k_scale_values(double* __restrict__ nz, const int64_t* __restrict__ idx,
                               double sc, int64_t n, int64_t base) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) nz[idx[i] - base] *= sc;
}

// ------------------------------------------------------------------ G1 cone -> K values
// Diagonal Hs entries (Zero, NN, sparse SOC): nz[map] = -Hs.
//   kind 0: 0 ; 1: w^2 (NN, coneops_nncone.jl:91-101) ; 2: eta^2 (sparse SOC tail) ;
//   3: eta^2 * d (sparse SOC head, coneops_socone.jl:161-166)
__global__ void k_hs_diag(int64_t n, const int8_t* __restrict__ kind, const int32_t* __restrict__ midx,
                          const int32_t* __restrict__ cone, const int64_t* __restrict__ map,
                          const double* __restrict__ w, const double* __restrict__ eta,
                          const double* __restrict__ dd, double* __restrict__ nz) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = kind[i];
    double v = 0.0;
    if (k == 1) { const double t = w[midx[i]]; v = t * t; }
    else if (k >= 2) { const double e = eta[cone[i]]; v = e * e; if (k == 3) v *= dd[cone[i]]; }
    nz[map[i]] = -v;
}

// Dense SOC blocks (dim <= 4): packed triu of eta^2 (2ww' - J), Hs[0] = (sqrt2 w0 - 1)(sqrt2 w0 + 1)
// (coneops_socone.jl:168-187).  One thread per cone.
__global__ void k_hs_soc_dense(int32_t ncone, const int32_t* __restrict__ moff,
                               const int32_t* __restrict__ dim, const int32_t* __restrict__ socid,
                               const int64_t* __restrict__ hoff, const int64_t* __restrict__ map,
                               const double* __restrict__ w, const double* __restrict__ eta,
                               double* __restrict__ nz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncone) return;
    const double* wc = w + moff[i];
    const int d = dim[i];
    const double e = eta[socid[i]];
    const double e2 = e * e;
    const int64_t* mp = map + hoff[i];
    const double s2 = sqrt(2.0);
    // explicit _rn intrinsics: no FMA contraction, so the values are bit-identical to the
    // reference's scalar arithmetic
    const double sw = __dmul_rn(s2, wc[0]);
    nz[mp[0]] = -__dmul_rn(__dmul_rn(__dadd_rn(sw, -1.0), __dadd_rn(sw, 1.0)), e2);
    int h = 1;
    for (int col = 1; col < d; ++col)
        for (int row = 0; row <= col; ++row) {
            double v = __dmul_rn(__dmul_rn(2.0, wc[row]), wc[col]);
            if (row == col) v = __dadd_rn(v, 1.0);
            nz[mp[h++]] = -__dmul_rn(v, e2);
        }
}

// Sparse SOC expansion columns: nz[map_u] = u * (-eta^2), nz[map_v] = v * (-eta^2)
// (directldl_datamaps.jl:61-79: update then scale by -eta^2), D = (-eta^2, +eta^2).
__global__ void k_soc_expansion(int64_t n, const int32_t* __restrict__ src, const int32_t* __restrict__ cone,
                                const int64_t* __restrict__ mapu, const int64_t* __restrict__ mapv,
                                const double* __restrict__ u, const double* __restrict__ v,
                                const double* __restrict__ eta, double* __restrict__ nz) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double e = eta[cone[i]];
    const double me2 = -(e * e);
    nz[mapu[i]] = u[src[i]] * me2;
    nz[mapv[i]] = v[src[i]] * me2;
}
__global__ void k_soc_D(int32_t n, const int32_t* __restrict__ cone, const int64_t* __restrict__ mapD,
                        const double* __restrict__ eta, double* __restrict__ nz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double e = eta[cone[i]];
    nz[mapD[2 * i]] = -(e * e);
    nz[mapD[2 * i + 1]] = e * e;
}

// ------------------------------------------------------------------ G2 PSD: Hs = (RR') (x)_s (RR')
// k_psd_rrt: A = R R' per cone (n x n, column-major), one CTA per cone.
__global__ void __launch_bounds__(256)
k_psd_rrt(const int32_t* __restrict__ side, const int64_t* __restrict__ roff,
          const double* __restrict__ R, double* __restrict__ A) {
    const int c = blockIdx.x;
    const int n = side[c];
    const double* Rc = R + roff[c];
    double* Ac = A + roff[c];
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e % n, j = e / n;
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += Rc[i + k * n] * Rc[j + k * n];
        Ac[e] = acc;
    }
}
// k_psd_skron: grid (column tiles, cones).  Column q=(k,l), row p=(i,j) of triu(A (x)_s A)
// (skron!, coneops_psdtrianglecone.jl:502-540); writes -value through the Hs map.
__global__ void __launch_bounds__(256)
k_psd_skron(const int32_t* __restrict__ side, const int64_t* __restrict__ roff,
            const int64_t* __restrict__ hoff, const double* __restrict__ A,
            const int64_t* __restrict__ map, double* __restrict__ nz) {
    extern __shared__ double sA[];
    const int c = blockIdx.y;
    const int n = side[c];
    const int ne = n * (n + 1) / 2;
    const double* Ac = A + roff[c];
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) sA[e] = Ac[e];
    __syncthreads();
    const int64_t* mp = map + hoff[c];
    const double s2 = sqrt(2.0);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int q = blockIdx.x * 8 + wid; q < ne; q += gridDim.x * 8) {
        // column q -> (k <= l)
        int l = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
        while ((l + 1) * (l + 2) / 2 <= q) ++l;
        while (l * (l + 1) / 2 > q) --l;
        const int k = q - l * (l + 1) / 2;
        const bool kl = (k == l);
        const int64_t cbase = (int64_t)q * (q + 1) / 2;
        // rows p = 0..q, p -> (i <= j)
        for (int p = lane; p <= q; p += 32) {
            int j = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
            while ((j + 1) * (j + 2) / 2 <= p) ++j;
            while (j * (j + 1) / 2 > p) --j;
            const int i = p - j * (j + 1) / 2;
            const bool ij = (i == j);
            const double Ajl = sA[j + l * n], Ajk = sA[j + k * n];
            double v;
            if (!ij && !kl) v = sA[i + k * n] * Ajl + sA[i + l * n] * Ajk;
            else if (ij && !kl) v = s2 * Ajl * Ajk;
            else if (!ij && kl) v = s2 * sA[i + l * n] * Ajk;
            else v = Ajl * Ajl;
            nz[mp[cbase + p]] = -v;
        }
    }
}

}  // namespace cb200
